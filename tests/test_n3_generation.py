"""N3: the generation config with `update_from_string`, the eval-batch rescale of do_evaluate, n-best
generation and its dump -- against goldens produced by the reference's own functions (tests/golden/make_golden_n3.py imports
src/decoding/config.py and src/utilities/generation_utils.py unmodified) and, where /root/reference is present, against
the live reference classes."""
import json
import os
import sys

import pytest
import torch

import parity
from hf_stub import StubConfig, StubDecoder
from huggingface_asr_b200.beam_search import joint_beam_search
from huggingface_asr_b200.generation import (GenerationConfigCustom, generate_nbest, joint_ctc_generation_config, override_for_evaluation,
                                             rescale_eval_batch, save_nbests)
from huggingface_asr_b200.synthetic import BLANK, BOS, EOS, make_attention_scores
from oracle import oracle as orc

G = json.load(open(os.path.join(parity.GOLDEN, "n3_generation.json")))
REF_SRC = "/root/reference/src"


def _apply(cls, base, update):
    c = cls(**base)
    err = None
    try:
        c.update_from_string(update)
    except Exception as e:  # noqa: BLE001
        err = type(e).__name__
    return c, err


@pytest.mark.parametrize("case", G["updates"], ids=[u["update"] for u in G["updates"]])
def test_update_from_string_vs_reference_golden(case):
    c, err = _apply(GenerationConfigCustom, case["base"], case["update"])
    assert err == case.get("error")
    assert {k: getattr(c, k) for k in case["result"]} == case["result"]
    for k, v in case["result"].items():
        assert type(getattr(c, k)) is type(v), f"{k}: {type(getattr(c, k)).__name__} vs {type(v).__name__}"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference is only present in the dev container")
def test_update_from_string_vs_live_reference():
    sys.path.insert(0, REF_SRC)
    try:
        from decoding.config import GenerationConfigCustom as Ref
    finally:
        sys.path.remove(REF_SRC)
    base = dict(ctc_weight=0.2, num_beams=5, apply_eos_space_trick=True, eos_space_trick_weight=1.25, max_length=40)
    for upd in ["num_beams=12;ctc_weight=0.7", "apply_eos_space_trick=No", "max_length=9;bogus=1", "eos_space_trick_weight=3", "x", "do_sample=TRUE"]:
        (a, ea), (b, eb) = _apply(GenerationConfigCustom, base, upd), _apply(Ref, base, upd)
        assert ea == eb, upd
        for k in ("ctc_weight", "num_beams", "apply_eos_space_trick", "eos_space_trick_weight", "max_length", "do_sample", "ctc_margin", "lm_weight"):
            assert getattr(a, k) == getattr(b, k), (upd, k)


def test_eval_batch_rescale_and_override():
    for r in G["rescale"]:
        assert rescale_eval_batch(r["eval_batch"], r["beams_orig"], r["beams_new"]) == r["result"]
    cfg = joint_ctc_generation_config(ctc_weight=0.3, num_beams=4, max_length=32)
    assert override_for_evaluation(cfg, None, 32) == 32 and cfg.num_beams == 4
    assert override_for_evaluation(cfg, "ctc_weight=0.5;num_beams=10", 32) == 13  # ceil(32 / 2.5), general_utils.py:144-147
    assert cfg.num_beams == 10 and cfg.ctc_weight == 0.5
    with pytest.raises(ValueError):
        joint_ctc_generation_config(ctc_pre_beam_size=1)


def test_save_nbests_files_vs_reference_golden(tmp_path):
    nb = G["nbests"]
    decode = lambda ids: " ".join(f"t{i}" for i in ids if i not in (0, 1, 2, 3))  # noqa: E731  the golden's tokenizer
    path = str(tmp_path / "nb")
    labels = [torch.tensor(x) for x in nb["labels"]]
    save_nbests(path, [torch.tensor(x) for x in nb["nbests"]], [torch.tensor(x) for x in nb["scores"]], labels, decode, 3, nb["group_size"])
    for suffix, want in nb["files"].items():
        assert open(path + suffix).read() == want, suffix
    assert (labels[0] == -100).any(), "the caller's labels must not be modified"


class OracleLogSoftmax:
    def __call__(self, input_ids, scores):
        return torch.log_softmax(scores, dim=-1)


class OracleStub(StubDecoder):
    ctc_rescorer_cls = orc.OracleCTCRescorerLogitsProcessor
    log_softmax_cls = OracleLogSoftmax


def _golden_decode(i):
    g = parity.load("decode_1best")
    return (torch.from_numpy(g[f"d{i}_logits"]), torch.from_numpy(g[f"d{i}_lens"]), int(g[f"d{i}_W"]), int(g[f"d{i}_seed"]),
            int(g[f"d{i}_max_length"]), g[f"d{i}_seq"], g[f"d{i}_len"])


def test_generate_nbest_under_hf_with_the_oracle_processor():
    """do_generate's loop: R hypotheses per utterance with scores, grouped utterance-major; rank 1 is the golden 1-best."""
    logits, lens, W, seed, ml, seq, ln = _golden_decode(1)
    B, T, V = logits.shape
    R = min(3, W)
    m = OracleStub(StubConfig(V), seed=seed)
    cfg = joint_ctc_generation_config(ctc_weight=0.3, num_beams=W, max_length=ml, pad_token_id=BLANK, eos_token_id=EOS, bos_token_id=BOS,
                                      do_sample=False, length_penalty=1.0, early_stopping=False, use_cache=True)
    labels = torch.from_numpy(seq[:, : int(ln.max())].copy())
    batch = {"inputs": torch.full((B, 1), BOS, dtype=torch.long), "labels": labels, "encoder_logits": logits, "encoder_output_lens": lens}
    out = generate_nbest(m, [batch], cfg, num_predictions_to_return=R, eval_beam_factor=1, eval_batch_size=8)
    assert out.group_size == R and out.eval_batch_size == 8 and cfg.num_return_sequences == R and cfg.output_scores
    seqs, scores = out.nbests[0], out.scores[0]
    assert seqs.shape[0] == B * R and scores.shape == (B * R,)
    sc = scores.view(B, R)
    assert (sc[:, :-1] >= sc[:, 1:]).all(), "hypotheses of an utterance must come best first"
    for b in range(B):
        n = int(ln[b])
        assert (seqs[b * R, 1: 1 + n].numpy() == seq[b, :n]).all()
    out2 = generate_nbest(OracleStub(StubConfig(V), seed=seed), [batch], joint_ctc_generation_config(
        ctc_weight=0.3, num_beams=W, max_length=ml, pad_token_id=BLANK, eos_token_id=EOS, bos_token_id=BOS, do_sample=False, length_penalty=1.0,
        early_stopping=False, use_cache=True), num_predictions_to_return=1, eval_beam_factor=2, eval_batch_size=7)
    assert out2.eval_batch_size == 4 and out2.nbests[0].shape[0] == B  # general_utils.py:199-203: beams doubled, batch halved (ceil)


def test_nbest_of_the_shared_harness():
    logits, lens, W, seed, ml, seq, ln = _golden_decode(1)
    B, T, V = logits.shape
    dec = lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5)  # noqa: E731
    R = min(3, W)
    a = joint_beam_search(orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W), dec, B, W, V, BOS, EOS,
                          BLANK, max_length=ml, num_return_sequences=R)
    assert a.nbest_sequences.shape == (B, R, ml) and a.nbest_scores.shape == (B, R)
    assert torch.equal(a.nbest_sequences[:, 0], a.sequences) and torch.equal(a.nbest_scores[:, 0], a.scores)
    assert (a.nbest_scores[:, :-1] >= a.nbest_scores[:, 1:]).all()
    for b in range(B):
        assert len({tuple(a.nbest_sequences[b, r].tolist()) for r in range(R) if a.nbest_scores[b, r] > float("-inf")}) >= 1
    with pytest.raises(ValueError):
        joint_beam_search(orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W), dec, B, W, V, BOS, EOS,
                          BLANK, max_length=4, num_return_sequences=W + 1)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("pre_beam", [0, 12])
def test_nbest_native_and_fused_loops_equal_the_torch_harness(pre_beam):
    from huggingface_asr_b200.beam_search import joint_beam_search_fused, joint_beam_search_native
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

    logits, lens, W, seed, ml, _, _ = _golden_decode(1)
    B, T, V = logits.shape
    R = min(4, W)
    dec = lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5).cuda()  # noqa: E731
    outs = []
    for loop in (joint_beam_search, joint_beam_search_fused, joint_beam_search_native):
        proc = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False,
                                          pre_beam_size=pre_beam)
        kw = {"done_check_lag": 0} if loop is not joint_beam_search else {}
        outs.append(loop(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=ml, device=torch.device("cuda"), num_return_sequences=R, **kw))
    a = outs[0]
    assert a.nbest_sequences.shape == (B, R, ml)
    for o in outs[1:]:
        assert torch.equal(a.nbest_sequences, o.nbest_sequences) and torch.equal(a.nbest_lengths, o.nbest_lengths)
        assert torch.equal(a.nbest_scores, o.nbest_scores)


@pytest.mark.gpu
def test_hf_generate_pre_beam_without_a_kv_cache_recovers_the_parents():
    """with use_cache=False transformers never calls _reorder_cache, so nobody hands the processor beam indices.  In
    pre-beam mode a token is only scored for the hypothesis that proposed it; the processor recovers every row's parent from
    the prefixes (or would have to refuse) -- the decode must equal the cached one, where the beam indices are reported."""
    from test_generation_hf import _generate

    logits, lens, W, seed, ml, _, _ = _golden_decode(1)
    a = _generate(StubDecoder, "cuda", logits, lens, W, seed, ml, True, ctc_pre_beam_size=12)
    b = _generate(StubDecoder, "cuda", logits, lens, W, seed, ml, False, ctc_pre_beam_size=12)
    assert a.shape == b.shape and (a == b).all()
