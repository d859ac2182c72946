"""GPU: the fused beam-search step (ctcps_beam_step) against the torch restatement of the HF loop."""
import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert a.steps == b.steps, f"steps {a.steps} vs {b.steps}"
    assert torch.equal(a.lengths, b.lengths), "lengths differ"
    assert torch.equal(a.sequences, b.sequences), "1-best token sequences differ"
    assert torch.equal(a.scores, b.scores), f"scores differ by {(a.scores - b.scores).abs().max().item()}"


@pytest.mark.parametrize("B,W,V,eos_boost,max_length,seed", [
    (5, 4, 50, 3.0, 24, 0),      # small vocabulary: eos everywhere, heavy pool traffic
    (3, 10, 333, 4.0, 40, 1),
    (2, 20, 97, 2.0, 30, 2),     # 2W = 40 > 32: two list entries per lane
    (4, 1, 64, 3.0, 20, 3),      # greedy width
    (6, 3, 1000, 6.0, 16, 4),    # runs into max_length with open beams
    (2, 32, 70, 3.0, 12, 5),     # widest supported beam
])
def test_fused_beam_search_equals_torch_harness_on_random_scores(B, W, V, eos_boost, max_length, seed):
    """Identity processor + random decoder: stresses top-2W, eos finalisation, pool replacement, done test, reordering."""
    from huggingface_asr_b200.beam_search import joint_beam_search, joint_beam_search_fused

    g = torch.Generator().manual_seed(seed)
    table = torch.randn(max_length + 1, B * W, V, generator=g)
    table[:, :, 1] += eos_boost * torch.rand(max_length + 1, B * W, generator=g)
    table = torch.log_softmax(table, -1).cuda()

    def decoder(ids, n):
        # depends on the row's last token so that reordering matters
        return table[n] + 0.01 * (ids[:, -1:] % 7).float()

    ident = lambda ids, s: s  # noqa: E731
    a = joint_beam_search(ident, decoder, B, W, V, 0, 1, 3, max_length=max_length, device="cuda")
    b = joint_beam_search_fused(ident, decoder, B, W, V, 0, 1, 3, max_length=max_length, device="cuda")
    _same(a, b)
    c = joint_beam_search_fused(ident, decoder, B, W, V, 0, 1, 3, max_length=max_length, device="cuda", done_check_lag=2)
    assert torch.equal(a.sequences, c.sequences) and torch.equal(a.scores, c.scores) and a.steps <= c.steps <= a.steps + 2


@pytest.mark.parametrize("materialize", [True, False])
def test_fused_decode_with_scorer_equals_torch_harness(materialize):
    from huggingface_asr_b200.beam_search import joint_beam_search, joint_beam_search_fused
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import BLANK, BOS, EOS, SyntheticDecoder, make_encoder_logits

    B, W, T, V = 6, 10, 96, 1000
    logits, lens, tr = make_encoder_logits(B, T, V, "peaky", True, seed=77)
    dec = SyntheticDecoder(tr, W, V, 40, seed=3, device="cuda")
    mk = lambda: CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0,  # noqa: E731
                                            materialize_state=materialize)
    a = joint_beam_search(mk(), dec, B, W, V, BOS, EOS, BLANK, max_length=40, device="cuda")
    b = joint_beam_search_fused(mk(), dec, B, W, V, BOS, EOS, BLANK, max_length=40, device="cuda")
    _same(a, b)
    for i in range(B):  # and the decode is right: the 1-best is the aligned transcript
        assert a.sequences[i, : a.lengths[i]].tolist() == tr[i][:-1]


def test_fused_decode_reproduces_reference_golden_1best():
    """The reference processor's 1-best (golden, torch harness on CPU) through CUDA scorer + fused beam step."""
    from huggingface_asr_b200.beam_search import joint_beam_search_fused
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores

    g = parity.load("decode_1best")
    for i in range(3):
        logits, lens = torch.from_numpy(g[f"d{i}_logits"]).cuda(), torch.from_numpy(g[f"d{i}_lens"]).cuda()
        W, seed = int(g[f"d{i}_W"]), int(g[f"d{i}_seed"])
        B, T, V = logits.shape
        for mat in (True, False):
            proc = CTCRescorerLogitsProcessor(logits.clone(), lens, 3, 1, 0, 0.3, W, -1, False, 1.0, materialize_state=mat)
            out = joint_beam_search_fused(proc, lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5).cuda(), B, W, V,
                                          0, 1, 3, max_length=int(g[f"d{i}_max_length"]), device="cuda")
            assert out.steps == int(g[f"d{i}_steps"])
            assert (out.sequences.cpu().numpy() == g[f"d{i}_seq"]).all(), f"decode {i}: 1-best differs from the reference"
            assert (out.lengths.cpu().numpy() == g[f"d{i}_len"]).all()
            assert np.abs(out.scores.cpu().numpy() - g[f"d{i}_score"]).max() <= 1e-4
