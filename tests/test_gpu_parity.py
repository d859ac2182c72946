"""GPU parity: the sm_100a scorer, called through the C ABI, against the golden vectors of the
reference and against the CPU oracle on seeded inputs.  Run on the B200 box with -m gpu."""
import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


class CudaBackend(parity.Backend):
    device = "cuda"

    def make_scorer(self, x_logp, lens, blank, eos, margin=0):
        from huggingface_asr_b200.decoding.ctc_scorer import CTCPrefixScoreTH

        return CTCPrefixScoreTH(x_logp.contiguous(), lens, blank, eos, margin)

    def make_processor(self, logits, lens, pad, eos, margin, w, W, space=-1, trick=False, trick_w=1.0):
        from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

        return CTCRescorerLogitsProcessor(logits, lens, pad, eos, margin, w, W, space, trick, trick_w, materialize_state=True)


class CudaLazyBackend(CudaBackend):
    """Same scorer in lazy-state mode: r is never materialised, survivors are recomputed at select time."""

    def make_scorer(self, x_logp, lens, blank, eos, margin=0):
        sc = super().make_scorer(x_logp, lens, blank, eos, margin)
        sc.lazy_state = True
        return sc

    def make_processor(self, logits, lens, pad, eos, margin, w, W, space=-1, trick=False, trick_w=1.0):
        from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

        return CTCRescorerLogitsProcessor(logits, lens, pad, eos, margin, w, W, space, trick, trick_w, materialize_state=False)


BE = CudaBackend()
BE_LAZY = CudaLazyBackend()
STEP_CASES = ["steps_peaky_w3", "steps_peaky_ragged_w10", "steps_flat_w1", "steps_flat_w20", "steps_peaky_w5_v129",
              "steps_forced_pad", "steps_trick"]


def test_native_library_is_loaded():
    from huggingface_asr_b200 import _lib

    assert _lib.lib().ctcps_version() >= 100


@pytest.mark.parametrize("name", STEP_CASES)
def test_steps_vs_reference_golden(name):
    worst = parity.replay_steps(BE, name)
    print(name, worst)


@pytest.mark.parametrize("name", STEP_CASES)
def test_lazy_steps_vs_reference_golden(name):
    worst = parity.replay_steps(BE_LAZY, name)
    print(name, worst)


def test_lazy_edges_and_decode_vs_reference_golden():
    parity.replay_edges(BE_LAZY)
    parity.replay_decode(BE_LAZY)


@pytest.fixture
def restore_select_mode():
    prev = parity.select_mode(-1)
    yield
    parity.select_mode(prev)


@pytest.mark.parametrize("pscan", [0, 1])
@pytest.mark.parametrize("B,W,T,V", [(3, 10, 100, 1200), (2, 7, 61, 517), (2, 20, 90, 260), (3, 1, 50, 300)])
def test_lazy_state_is_bit_identical_to_materialised(B, W, T, V, pscan, restore_select_mode):
    """Lazy mode must reproduce the materialising kernels: same joint scores, and selected states that are bit-identical
    with the sequential selection kernel (same operations in the same order per lane) and inside the parity criterion
    with the time-parallel one (the library default; a different fp32 evaluation order, adjudicated against fp64 in
    test_gpu_select_pscan.py)."""
    parity.select_mode(pscan)
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor, LazyForwardVariables
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits

    logits, lens, _ = make_encoder_logits(B, T, V, "peaky", True, seed=4321 + W)
    mat = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0, materialize_state=True)
    lazy = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
    ids = torch.zeros((B * W, 1), dtype=torch.long, device="cuda")
    beam_scores = torch.zeros(B, W, device="cuda")
    beam_scores[:, 1:] = -1e9
    for n in range(6):
        att = make_attention_scores(B * W, V, n, seed=5, scale=0.5).cuda()
        out_m = mat(ids, att.clone())
        out_l = lazy(ids, att.clone())
        assert isinstance(lazy.ctc_states[0], LazyForwardVariables)
        # log_psi: the linear-domain sum is accumulated in a different association (all hyps of a thread) -> ulp-level
        assert (out_m - out_l).abs().max().item() <= 2e-5, f"step {n}: joint scores differ"
        if n > 0:
            pass
        r_l = lazy.ctc_states[0].materialize()
        if pscan == 0:
            assert torch.equal(r_l, mat.ctc_states[0]), f"step {n}: materialised lazy state differs"
        else:  # the lazy chain of selected states went through the time-parallel kernel: ulp-level differences propagate
            parity.assert_parity(r_l, mat.ctc_states[0], f"step {n}: materialised lazy state (time-parallel selection chain)")
        sel_m = mat.ctc_prefix_scorer.index_select_state(mat.ctc_states, ids[:, -1].reshape(-1, W) * 0 + 5)
        sel_l = lazy.ctc_prefix_scorer.index_select_state(lazy.ctc_states, ids[:, -1].reshape(-1, W) * 0 + 5)
        if pscan == 0:
            assert torch.equal(sel_m[0], sel_l[0]), f"step {n}: recomputed survivors differ from the gathered ones"
        else:
            parity.assert_parity(sel_l[0], sel_m[0], f"step {n}: time-parallel survivors vs the gathered ones")
        cand = (out_m + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        base = (torch.arange(B, device="cuda") * W).view(B, 1)
        ids = torch.cat([ids[(src + base).view(-1)], tok.view(-1, 1)], dim=1)
        beam_scores = top


def test_partial_scoring_vs_reference_golden():
    parity.replay_partial(BE)


def test_select_general_vs_reference_golden():
    parity.replay_select_general(BE)


def test_edges_vs_reference_golden():
    parity.replay_edges(BE)


@pytest.mark.parametrize("lazy", [False, True])
@pytest.mark.parametrize("name", parity.WINDOW_CASES)
def test_attention_window_vs_reference_golden(name, lazy):
    """margin > 0 with attention weights (ctc_scorer.py:127-136): full vocabulary and scoring_ids, on a materialising scorer
    and on a lazy one (windowed calls go through the materialising kernels either way)."""
    print(name, parity.replay_window(BE_LAZY if lazy else BE, name))


def test_extend_prob_and_state_vs_reference_golden():
    parity.replay_extend(BE)
    parity.replay_extend(BE_LAZY)


def test_processor_input_validation_and_strided_scores():
    """Wrong vocabulary raises; a non-contiguous `scores` tensor still gets its pad column set in place (:325)."""
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits

    B, W, T, V = 2, 3, 20, 32
    logits, lens, _ = make_encoder_logits(B, T, V, "flat", True, seed=8)
    ids = torch.zeros((B * W, 1), dtype=torch.long, device="cuda")
    att = make_attention_scores(B * W, V, 0, seed=8).cuda()
    ref = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0)(ids, att.clone())
    wide = torch.zeros(B * W, 2 * V, device="cuda")
    wide[:, ::2] = att
    strided = wide[:, ::2]
    assert not strided.is_contiguous()
    out = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0)(ids, strided)
    assert torch.equal(out, ref)
    assert (strided[:, 3] == -1e10).all() and (wide[:, 6] == -1e10).all()
    with pytest.raises(ValueError, match="vocabulary"):
        CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0)(ids, torch.zeros(B * W, V + 1, device="cuda"))
    with pytest.raises(ValueError, match="multiple of the batch"):
        CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0)(ids[:5], att[:5].clone())
    # half-precision encoder outputs (autocast) are upcast like the reference's log_softmax would accept them; integers raise
    half = CTCRescorerLogitsProcessor(logits.cuda().half(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0)(ids, att.clone())
    same = CTCRescorerLogitsProcessor(logits.cuda().half().float(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0)(ids, att.clone())
    assert torch.equal(half, same)
    with pytest.raises(ValueError, match="float32"):
        CTCRescorerLogitsProcessor(logits.cuda().long(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0)
    with pytest.raises(ValueError, match="pre_beam_size"):
        CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0, pre_beam_size=1)


def test_decode_1best_vs_reference_golden():
    parity.replay_decode(BE)


def test_padded_posteriors_vs_reference_golden():
    from huggingface_asr_b200.decoding.ctc_scorer import CTCPrefixScoreTH

    g = parity.load("steps_peaky_ragged_w10")
    sc = CTCPrefixScoreTH.from_logits(BE.t(g["logits"]), BE.t(g["lens"]), 3, 1)
    parity.assert_parity(sc._x[:, :, : sc.odim], g["x_padded"], "padded log-posteriors", atol=2e-6, rtol=0)
    ref_x = torch.from_numpy(g["x_padded"]).transpose(0, 1)
    parity.assert_parity(sc.x[0], ref_x, "x property plane 0", atol=2e-6, rtol=0)
    parity.assert_parity(sc.x[1], ref_x[:, :, 3:4].expand(-1, -1, sc.odim), "x property plane 1", atol=2e-6, rtol=0)


@pytest.mark.parametrize("B,W,T,V,kind,ragged", [
    (4, 10, 120, 1000, "peaky", True),     # several v-tiles, ragged, two hyp groups
    (2, 7, 61, 517, "flat", True),         # W with a padded hyp group, V % 4 != 0, V tail inside a TMA box
    (3, 1, 90, 300, "peaky", False),       # greedy-width
    (2, 20, 200, 260, "flat", False),      # W = 20 (four groups)
])
def test_multi_step_vs_oracle(B, W, T, V, kind, ragged):
    """Seeded inputs at sizes the oracle finishes in seconds; every step of a short decode, r included."""
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits
    from oracle import oracle as orc

    logits, lens, _ = make_encoder_logits(B, T, V, kind, ragged, seed=1234 + W)
    gpu = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0, materialize_state=True)
    cpu = orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), 3, 1, 0, 0.3, W)
    cpu64 = orc.OracleCTCRescorerLogitsProcessor(logits.double(), lens.clone(), 3, 1, 0, 0.3, W)
    ids = torch.zeros((B * W, 1), dtype=torch.long)
    beam_scores = torch.zeros(B, W)
    beam_scores[:, 1:] = -1e9
    for n in range(5):
        att = make_attention_scores(B * W, V, n, seed=99, scale=0.5)
        out_c = cpu(ids, att.clone())
        out_64 = cpu64(ids, att.double())
        out_g = gpu(ids.cuda(), att.cuda())
        parity.assert_parity(out_g, out_c, f"step {n} joint scores", ref64=out_64)
        parity.assert_parity(gpu.ctc_states[1], cpu.ctc_states[1], f"step {n} log_psi", ref64=cpu64.ctc_states[1])
        parity.assert_parity(gpu.ctc_states[0], cpu.ctc_states[0], f"step {n} r", ref64=cpu64.ctc_states[0])
        cand = (out_c + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        ids = torch.cat([ids[(src + (torch.arange(B) * W).view(B, 1)).view(-1)], tok.view(-1, 1)], dim=1)
        beam_scores = top


def test_log_softmax_processor():
    from huggingface_asr_b200.decoding.ctc_scorer import LogSoftmaxProcessor

    g = torch.Generator().manual_seed(3)
    s = torch.randn(7, 5000, generator=g) * 3
    out = LogSoftmaxProcessor()(None, s.cuda())
    assert (out.cpu() - torch.log_softmax(s, -1)).abs().max() <= 2e-6


def test_cpu_tensors_are_refused():
    from huggingface_asr_b200.decoding.ctc_scorer import CTCPrefixScoreTH

    with pytest.raises(RuntimeError, match="no CPU path"):
        CTCPrefixScoreTH(torch.zeros(1, 4, 8), torch.tensor([4]), 3, 1)
