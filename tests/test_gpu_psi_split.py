"""GPU: the frame-split lazy scoring kernel (k_psi_split) against the one-CTA-per-tile kernel (k_psi_full) and the oracle.

Shapes are chosen so that a tile is shared by many CTAs (few tiles, long T), by exactly two (tiles ~ resident CTAs), and
by one (more tiles than CTAs), with one and several hypothesis groups, V % 4 != 0, ragged lengths, first step (ol = 0:
the first-frame term is read by whichever CTA finishes the tile) and later steps (a shorter frame range).
"""
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _mode(m):
    from huggingface_asr_b200 import _lib

    return _lib.lib().ctcps_set_psi_split(m)


@pytest.fixture(autouse=True)
def _restore_mode():
    prev = _mode(-1)
    yield
    _mode(prev)


def _decode_steps(B, W, T, V, kind, ragged, steps, seed):
    """Three lazy processors in lockstep on the same hypotheses (those of the whole-tile kernel): whole tiles, split, split."""
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits

    logits, lens, _ = make_encoder_logits(B, T, V, kind, ragged, seed=seed)
    modes = (0, 1, 1)
    procs = [CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
             for _ in modes]
    ids = torch.zeros((B * W, 1), dtype=torch.long, device="cuda")
    beam_scores = torch.zeros(B, W, device="cuda")
    beam_scores[:, 1:] = -1e9
    outs = []
    for n in range(steps):
        att = make_attention_scores(B * W, V, n, seed=seed + 1, scale=0.5).cuda()
        step = []
        for m, proc in zip(modes, procs):
            _mode(m)
            out = proc(ids, att.clone())
            step.append((out.clone(), proc.ctc_states[1].clone()))
        outs.append(step)
        cand = (step[0][0] + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        base = (torch.arange(B, device="cuda") * W).view(B, 1)
        ids = torch.cat([ids[(src + base).view(-1)], tok.view(-1, 1)], dim=1)
        beam_scores = top
    torch.cuda.synchronize()
    return outs


SHAPES = [
    (1, 10, 373, 5000, "peaky", False),   # 10 tiles for ~592 CTAs: every tile is summed from many pieces
    (2, 20, 748, 1000, "peaky", True),    # two hypothesis groups, long T, ragged
    (16, 10, 248, 5000, "peaky", True),   # C1: 160 tiles, 3-5 pieces per tile
    (64, 10, 96, 5000, "flat", True),     # about as many tiles as CTAs: whole tiles and two-piece tiles mixed
    (3, 7, 61, 517, "flat", True),        # padded hypothesis group, V % 4 != 0, last V-tile mostly empty
    (2, 1, 40, 300, "peaky", False),      # greedy width, fewer chunks than PSI_MIN_Q per tile
    (150, 3, 50, 2100, "peaky", True),    # more tiles (750) than CTAs, short frame range
]


@pytest.mark.parametrize("B,W,T,V,kind,ragged", SHAPES)
def test_split_equals_whole_tiles(B, W, T, V, kind, ragged):
    outs = _decode_steps(B, W, T, V, kind, ragged, 4, seed=77 + W)
    for n, ((jw, pw), (js, ps), (ja, pa)) in enumerate(outs):
        # only the association of the fp32 sum over frames differs
        fin = pw > -1e9
        assert bool((ps[~fin] <= -1e9).all()), f"step {n}: logzero class differs"
        assert (ps[fin] - pw[fin]).abs().max().item() <= 2e-5, f"step {n}: log_psi differs"
        finj = jw > -1e9
        assert bool((js[~finj] <= -1e9).all()) and (js[finj] - jw[finj]).abs().max().item() <= 2e-5, f"step {n}: joint differs"
        # pieces are merged in chunk order whatever the arrival order: run to run bit-identical
        assert torch.equal(js, ja) and torch.equal(ps, pa), f"step {n}: the split kernel is not deterministic"


@pytest.mark.parametrize("B,W,T,V,kind,ragged", [(2, 10, 200, 1100, "peaky", True), (1, 20, 160, 520, "flat", False)])
def test_split_vs_oracle(B, W, T, V, kind, ragged):
    """Few tiles, many pieces per tile, against the CPU oracle (fp32, adjudicated by fp64), every step of a short decode."""
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits
    from oracle import oracle as orc

    assert _mode(1) in (0, 1)
    logits, lens, _ = make_encoder_logits(B, T, V, kind, ragged, seed=4242 + W)
    gpu = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), 3, 1, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
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
        cand = (out_c + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        ids = torch.cat([ids[(src + (torch.arange(B) * W).view(B, 1)).view(-1)], tok.view(-1, 1)], dim=1)
        beam_scores = top
