"""LM shallow fusion: LMRescorerLogitsProcessor with a KV cache against the golden outputs of the reference's
full-prefix processor (src/decoding/shallow_fussion.py:5-58; fixtures tests/golden/lm_fusion_*.npz made by
tests/golden/make_golden.py::case_lm_fusion with a small random GPT-2 whose weights are stored in the fixture)."""
import pytest
import torch

import parity
from hf_stub import StubConfig, StubDecoder
from huggingface_asr_b200.decoding.shallow_fusion import LMRescorerLogitsProcessor
from huggingface_asr_b200.generation import joint_ctc_generation_config
from huggingface_asr_b200.synthetic import BLANK, BOS, EOS

ATOL = 2e-5  # cached and full-prefix attention sum the same terms in a different order


def _lm(g):
    from transformers import GPT2Config, GPT2LMHeadModel

    cfg = GPT2Config(vocab_size=int(g["V"]), n_positions=64, n_embd=32, n_layer=2, n_head=4, bos_token_id=BOS, eos_token_id=EOS,
                     resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    lm = GPT2LMHeadModel(cfg).eval()
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("lm.")}
    lm.load_state_dict(sd)
    return lm


def _replay(name, use_cache=True, hand_beam_idx=False, device="cpu"):
    g = parity.load(name)
    proc = LMRescorerLogitsProcessor(float(g["lm_weight"]), _lm(g), torch.device(device), use_cache=use_cache)
    worst = 0.0
    for n in range(int(g["n_steps"])):
        ids = torch.from_numpy(g[f"input_ids_{n}"]).to(device)
        att = torch.from_numpy(g[f"att_{n}"]).to(device)
        if hand_beam_idx and n > 0:
            proc.set_beam_idx(torch.from_numpy(g[f"beam_idx_{n - 1}"]).to(device))
        out = proc(ids, att.clone()).cpu()
        ref = torch.from_numpy(g[f"out_{n}"])
        assert out.shape == ref.shape
        worst = max(worst, (out - ref).abs().max().item())
    return proc, worst, int(g["n_steps"])


@pytest.mark.parametrize("name", ["lm_fusion_w4", "lm_fusion_w1"])
@pytest.mark.parametrize("hand_beam_idx", [False, True])
def test_cached_lm_rescorer_vs_reference_golden(name, hand_beam_idx):
    proc, worst, n = _replay(name, True, hand_beam_idx)
    assert worst <= ATOL, f"{name}: cached LM scores differ from the reference's full-prefix scores by {worst}"
    # the LM saw the whole prefix once; every later step fed one token
    assert proc.full_forwards == 1 and proc.cached_forwards == n - 1


@pytest.mark.parametrize("name", ["lm_fusion_w4", "lm_fusion_w1"])
def test_uncached_lm_rescorer_is_the_reference(name):
    proc, worst, n = _replay(name, use_cache=False)
    assert worst <= 2e-6 and proc.full_forwards == n and proc.cached_forwards == 0


def test_unmatched_prefix_falls_back_to_the_full_forward():
    """A caller that edits a prefix (or hands a wrong beam_idx) must still get the full-prefix result."""
    g = parity.load("lm_fusion_w4")
    lm = _lm(g)
    proc = LMRescorerLogitsProcessor(0.5, lm, torch.device("cpu"))
    ref = LMRescorerLogitsProcessor(0.5, lm, torch.device("cpu"), use_cache=False)
    ids3 = torch.from_numpy(g["input_ids_3"])
    ids4 = torch.from_numpy(g["input_ids_4"]).clone()
    att = torch.from_numpy(g["att_4"])
    proc(ids3, att.clone())
    proc.set_beam_idx(torch.zeros(ids4.shape[0], dtype=torch.long))  # wrong on purpose: verified against the prefixes, ignored
    out = proc(ids4, att.clone())
    assert (out - ref(ids4, att.clone())).abs().max().item() <= ATOL and proc.cached_forwards == 1
    ids5 = torch.from_numpy(g["input_ids_5"]).clone()
    ids5[2, 1] = (ids5[2, 1] + 1) % int(g["V"])  # a prefix no cache row holds
    out = proc(ids5, att.clone())
    assert proc.full_forwards == 2
    assert (out - ref(ids5, att.clone())).abs().max().item() <= 2e-6
    # a new, shorter sequence (next generate()) restarts from a full forward as well
    out = proc(ids3, att.clone())
    assert proc.full_forwards == 3 and (out - ref(ids3, att.clone())).abs().max().item() <= 2e-6


class _UncachedLM(LMRescorerLogitsProcessor):
    def __init__(self, lm_weight, lm_model, device):
        super().__init__(lm_weight, lm_model, device, use_cache=False)


class _UncachedStub(StubDecoder):
    lm_rescorer_cls = _UncachedLM


@pytest.mark.parametrize("W,use_cache", [(4, True), (4, False), (1, True)])
def test_hf_generate_with_lm_fusion(W, use_cache):
    """transformers' own generate() with the mixin: lm_weight > 0 appends the processor (reference :398-404); with a KV
    cache HF's beam_idx reaches it through _reorder_cache; sequences equal those of the uncached (reference) processor."""
    g = parity.load("lm_fusion_w4")
    lm = _lm(g)
    V, B = int(g["V"]), 3
    cfg = joint_ctc_generation_config(lm_weight=0.5, num_beams=W, max_length=12, pad_token_id=BLANK, eos_token_id=EOS, bos_token_id=BOS,
                                      do_sample=False, length_penalty=1.0, early_stopping=False, use_cache=use_cache)
    outs = []
    for cls in (StubDecoder, _UncachedStub):
        m = cls(StubConfig(V), seed=21, raw_logits=False)
        m.set_lm_model(lm)
        seen = {}
        orig = m._get_logits_processor

        def spy(*a, _orig=orig, _m=m, _seen=seen, **k):
            procs = _orig(*a, **k)
            _seen["lm"] = _m.lm_rescorer
            return procs

        m._get_logits_processor = spy
        out = m.generate(torch.full((B, 1), BOS, dtype=torch.long), generation_config=cfg)
        assert m.lm_rescorer is None  # nothing of a batch survives generate()
        outs.append((out, seen["lm"]))
    assert torch.equal(outs[0][0], outs[1][0])
    cached = outs[0][1]
    assert cached.full_forwards == 1 and cached.cached_forwards >= 1
    assert outs[1][1].cached_forwards == 0


def test_lm_weight_without_a_model_raises():
    cfg = joint_ctc_generation_config(lm_weight=0.5, num_beams=2, max_length=6, pad_token_id=BLANK, eos_token_id=EOS, bos_token_id=BOS,
                                      do_sample=False)
    m = StubDecoder(StubConfig(16), seed=1)
    with pytest.raises(ValueError, match="lm_model"):
        m.generate(torch.full((2, 1), BOS, dtype=torch.long), generation_config=cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lm_fusion_w4", "lm_fusion_w1"])
def test_cached_lm_rescorer_on_the_gpu_vs_reference_golden(name):
    """Same replay with the LM and its cache on cuda:0 (the device the ASR model decodes on)."""
    proc, worst, n = _replay(name, True, hand_beam_idx=(name == "lm_fusion_w4"), device="cuda")
    assert worst <= 1e-4, f"{name}: {worst}"
    assert proc.full_forwards == 1 and proc.cached_forwards == n - 1
