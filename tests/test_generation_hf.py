"""N3: the scorer inside transformers' own `generate()` (HF 5.x beam search / greedy search), through the mixin that
restates JointCTCAttentionEncoderDecoder's hooks (ctc_encoder_plus_autoregressive_decoder.py:360-482).

CPU tests run the mixin with the oracle's processor class and pin HF's real beam search to the golden 1-best of the
reference processor; GPU tests run the sm_100a processor in the same loop and compare with the oracle run."""
import numpy as np
import pytest
import torch

import parity
from hf_stub import StubConfig, StubDecoder
from huggingface_asr_b200.generation import DecodingReport, edit_distance, error_rate, evaluate_decoding, joint_ctc_generation_config
from huggingface_asr_b200.synthetic import BLANK, BOS, EOS, make_attention_scores
from oracle import oracle as orc


class OracleLogSoftmax:
    def __call__(self, input_ids, scores):
        return torch.log_softmax(scores, dim=-1)


class OracleStub(StubDecoder):
    ctc_rescorer_cls = orc.OracleCTCRescorerLogitsProcessor
    log_softmax_cls = OracleLogSoftmax


def _cfg(W, max_length, use_cache, **kw):
    return joint_ctc_generation_config(ctc_weight=0.3, num_beams=W, max_length=max_length, pad_token_id=BLANK, eos_token_id=EOS,
                                       bos_token_id=BOS, do_sample=False, length_penalty=1.0, early_stopping=False,
                                       use_cache=use_cache, **kw)


def _golden_decode(i):
    g = parity.load("decode_1best")
    return (torch.from_numpy(g[f"d{i}_logits"]), torch.from_numpy(g[f"d{i}_lens"]), int(g[f"d{i}_W"]), int(g[f"d{i}_seed"]),
            int(g[f"d{i}_max_length"]), g[f"d{i}_seq"], g[f"d{i}_len"])


def _generate(cls, device, logits, lens, W, seed, max_length, use_cache, **cfg_kw):
    B, T, V = logits.shape
    m = cls(StubConfig(V), seed=seed, raw_logits=(W == 1)).to(device)
    m.set_ctc_inputs(logits.clone().to(device), lens.clone().to(device))
    out = m.generate(torch.full((B, 1), BOS, dtype=torch.long, device=device), generation_config=_cfg(W, max_length, use_cache, **cfg_kw))
    assert m.encoder_logits is None and m.ctc_rescorer is None  # nothing of a batch survives generate() (reference :480-481)
    return out[:, 1:].cpu()


@pytest.mark.parametrize("i", [0, 1, 2])
@pytest.mark.parametrize("use_cache", [False, True])
def test_hf_beam_search_with_oracle_processor_gives_the_reference_1best(i, use_cache):
    logits, lens, W, seed, ml, seq, ln = _golden_decode(i)
    out = _generate(OracleStub, "cpu", logits, lens, W, seed, ml, use_cache)
    for b in range(logits.shape[0]):
        n = int(ln[b])
        assert (out[b, :n].numpy() == seq[b, :n]).all(), f"decode {i} utterance {b}: HF beam search 1-best differs from the golden"


def test_hf_reorder_cache_hands_beam_idx_to_a_pre_beam_processor():
    """With a KV cache HF calls _reorder_cache(past, beam_idx); the mixin forwards it, so pre-beam decoding under HF selects
    states with hyp*V+tok like the shared harness does."""
    from huggingface_asr_b200.beam_search import joint_beam_search

    logits, lens, W, seed, ml, _, _ = _golden_decode(1)
    B, T, V = logits.shape
    out = _generate(OracleStub, "cpu", logits, lens, W, seed, ml, True, ctc_pre_beam_size=12)
    ref = joint_beam_search(orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W, pre_beam_size=12),
                            lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5), B, W, V, BOS, EOS, BLANK, max_length=ml)
    for b in range(B):
        n = int(ref.lengths[b])
        assert (out[b, :n] == ref.sequences[b, :n]).all()


def test_missing_ctc_inputs_raise():
    m = OracleStub(StubConfig(32))
    with pytest.raises(ValueError):
        m.generate(torch.full((1, 1), BOS, dtype=torch.long), generation_config=_cfg(2, 8, False))


def test_error_rate_and_report():
    assert edit_distance([1, 2, 3], [1, 2, 3]) == 0
    assert edit_distance([1, 2, 3], [1, 3]) == 1
    assert edit_distance([], [4, 5]) == 2
    assert edit_distance("kitten", "sitting") == 3
    assert error_rate([[1, 2, 3, 4], [5]], [[1, 2, 4], [5, 6]]) == pytest.approx(2 / 5)
    assert np.isnan(DecodingReport().tokens_per_second)


def test_evaluate_decoding_on_cpu_with_the_oracle():
    logits, lens, W, seed, ml, seq, ln = _golden_decode(0)
    B, T, V = logits.shape
    m = OracleStub(StubConfig(V), seed=seed)
    labels = torch.from_numpy(seq[:, : int(ln.max())].copy())
    labels[labels == BLANK] = -100
    batch = {"inputs": torch.full((B, 1), BOS, dtype=torch.long), "labels": labels, "encoder_logits": logits, "encoder_output_lens": lens}
    rep = evaluate_decoding(m, [batch, dict(batch)], _cfg(W, ml, False), pad_token_id=BLANK, special_token_ids=(BOS, EOS))
    assert rep.utterances == 2 * B and rep.error_rate == 0.0 and rep.tokens_produced > 0 and rep.tokens_per_second > 0
    assert rep.predictions[0] == [t for t in seq[0, : int(ln[0])].tolist() if t not in (BOS, EOS, BLANK)]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("i", [0, 1, 2])
@pytest.mark.parametrize("use_cache", [False, True])
def test_hf_beam_search_with_the_cuda_processor(i, use_cache):
    logits, lens, W, seed, ml, seq, ln = _golden_decode(i)
    out = _generate(StubDecoder, "cuda", logits, lens, W, seed, ml, use_cache)
    for b in range(logits.shape[0]):
        n = int(ln[b])
        assert (out[b, :n].numpy() == seq[b, :n]).all(), f"decode {i} utterance {b}: 1-best differs from the reference's golden"


@pytest.mark.gpu
@pytest.mark.parametrize("pre_beam", [0, 12])
def test_hf_generate_cuda_equals_oracle(pre_beam):
    """Beam search with and without pre-beam (KV cache on, so the beam_idx hook runs) and greedy search (LogSoftmaxProcessor)."""
    logits, lens, W, seed, ml, _, _ = _golden_decode(1)
    kw = {"ctc_pre_beam_size": pre_beam} if pre_beam else {}
    for beams in (W, 1):
        a = _generate(StubDecoder, "cuda", logits, lens, beams, seed, ml, True, **kw)
        b = _generate(OracleStub, "cpu", logits, lens, beams, seed, ml, True, **kw)
        assert a.shape == b.shape and (a == b).all(), f"num_beams={beams}, pre_beam={pre_beam}"


@pytest.mark.gpu
def test_evaluate_decoding_on_gpu():
    logits, lens, W, seed, ml, seq, ln = _golden_decode(0)
    B, T, V = logits.shape
    m = StubDecoder(StubConfig(V), seed=seed).cuda()
    labels = torch.from_numpy(seq[:, : int(ln.max())].copy())
    batch = {"inputs": torch.full((B, 1), BOS, dtype=torch.long, device="cuda"), "labels": labels, "encoder_logits": logits.cuda(),
             "encoder_output_lens": lens.cuda()}
    rep = evaluate_decoding(m, [batch], _cfg(W, ml, True), pad_token_id=BLANK, special_token_ids=(BOS, EOS))
    assert rep.error_rate == 0.0 and rep.utterances == B
