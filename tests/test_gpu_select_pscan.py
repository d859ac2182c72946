"""GPU: the time-parallel lazy state selection (k_select_lazy_pscan, the library default since round 2) against the
sequential kernel, the golden traces of the reference, the CPU oracle and -- as the judge of the two fp32 evaluation
orders -- the fp64 oracle.

History: written at the end of round 1, first hardware run failed the then-criterion at T = 748 (1.95e-3 at r = -872
between the two kernels).  Round 2 adjudicated with fp64 (this file): the sequential kernel itself is 1.7e-3 .. 2.2e-3
from fp64 on those entries (747 fp32 roundings at |r| ~ 900, one ulp = 6e-5), the time-parallel kernel is no further,
and joint scores / log_psi of both stay ~1e-6 from fp64.  22/22 passed on B200, the kernel
is faster on every BASELINE shape, and the default was flipped (CTCPS_SELECT_PSCAN=0 selects the sequential kernel).
"""
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu

BLANK, EOS, BOS = 3, 1, 0


def _mode(m):
    from huggingface_asr_b200 import _lib

    return _lib.lib().ctcps_set_select_pscan(m)


@pytest.fixture(autouse=True)
def _restore_mode():
    prev = _mode(-1)
    yield
    _mode(prev)


def _proc(logits, lens, W, w=0.3, **kw):
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

    return CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, w, W, -1, False, 1.0, materialize_state=False, **kw)


_f64_errors = parity.f64_errors


@pytest.mark.parametrize("B,W,T,V,kind", [
    (3, 10, 100, 1200, "peaky"),   # F = 4 frames per lane
    (2, 7, 61, 517, "flat"),       # fewer hypotheses than a CTA holds, V % 4 != 0
    (2, 20, 748, 260, "peaky"),    # C4 length: F = 24, shared-memory tile above 48 KB (the shape that failed the round-1 gate)
    (2, 20, 748, 260, "flat"),     # same length, every prefix plausible: the largest |r|
    (16, 10, 248, 1000, "peaky"),  # C1 length and batch, a vocabulary the fp64 oracle still holds in memory
    (16, 10, 248, 5000, "peaky"),  # C1 itself (too large for the fp64 oracle: kernel against kernel only)
    (1, 1, 40, 64, "flat"),        # one hypothesis, T - start barely above the warp width
    (5, 3, 20, 50, "peaky"),       # T < 32: most lanes have no frame
])
def test_pscan_equals_the_sequential_scan(B, W, T, V, kind):
    """Two lazy processors in lockstep on the same hypotheses; the selection inside __call__ runs sequentially in one and
    time-parallel in the other.  Both are two fp32 evaluation orders of the same recursion and neither is the reference, so
    the dispute is settled by a THIRD run: the fp64 oracle on the same inputs (CPU).  Against fp64
      * joint scores and log_psi: |time-parallel - fp64| <= |sequential - fp64| + 1e-4 element by element, and the
        time-parallel kernel itself stays inside the plain 1e-4 parity bound wherever the sequential one does;
      * selected forward variables: the time-parallel kernel's worst error is at most max(1e-4, 2 x the sequential kernel's worst error)
        (at |r| ~ 900 one fp32 ulp is 6e-5: an absolute 1e-4 is not representable, the sequential kernel is the yardstick).
    Where the fp64 oracle does not fit in memory the two kernels are compared with each other under the parity criterion."""
    import numpy as np

    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits
    from oracle import oracle as orc

    logits, lens, _ = make_encoder_logits(B, T, V, kind, True, seed=606 + W)
    procs = [_proc(logits.cuda(), lens.cuda(), W) for _ in range(2)]
    use64 = T * 2 * B * W * V * 8 <= 800 * 2**20
    p64 = orc.OracleCTCRescorerLogitsProcessor(logits.double(), lens, BLANK, EOS, 0, 0.3, W) if use64 else None
    ids = torch.zeros((B * W, 1), dtype=torch.long, device="cuda")
    beam_scores = torch.zeros(B, W, device="cuda")
    beam_scores[:, 1:] = -1e9
    worst = {"r_seq": 0.0, "r_par": 0.0, "joint_seq": 0.0, "joint_par": 0.0, "psi_seq": 0.0, "psi_par": 0.0}
    for n in range(min(7, T - 2)):
        att = make_attention_scores(B * W, V, n, seed=9, scale=0.5)
        outs, sels = [], []
        sel64 = out64 = None
        if use64:
            if n > 0:
                sel64 = p64.ctc_prefix_scorer.index_select_state(p64.ctc_states, ids[:, -1].reshape(-1, W).cpu())
            out64 = p64(ids.cpu(), att.double().clone())
        for m, proc in enumerate(procs):
            _mode(m)
            if n > 0:  # the same selection the processor is about to make, observed from outside
                sel = proc.ctc_prefix_scorer.index_select_state(proc.ctc_states, ids[:, -1].reshape(-1, W))
                sels.append((sel[0].clone(), sel[1].clone()))
            outs.append(proc(ids, att.cuda()).clone())
        if use64:
            if n > 0:
                e_seq, e_par = _f64_errors(sels[0][0], sel64[0]), _f64_errors(sels[1][0], sel64[0])
                worst["r_seq"], worst["r_par"] = max(worst["r_seq"], e_seq.max()), max(worst["r_par"], e_par.max())
                assert e_par.max() <= max(1e-4, 2 * e_seq.max()), (f"step {n} selected forward variables: time-parallel {e_par.max():.3e} vs "
                                                                    f"sequential {e_seq.max():.3e} from fp64")
                s_seq, s_par = _f64_errors(sels[0][1][:, 0], sel64[1][:, 0]), _f64_errors(sels[1][1][:, 0], sel64[1][:, 0])
                assert (s_par <= s_seq + 1e-4).all(), f"step {n} selected prefix scores: {s_par.max():.3e} vs {s_seq.max():.3e}"
            for key, a_seq, a_par, ref in (("joint", outs[0], outs[1], out64),
                                           ("psi", procs[0].ctc_states[1], procs[1].ctc_states[1], p64.ctc_states[1])):
                e_seq, e_par = _f64_errors(a_seq, ref), _f64_errors(a_par, ref)
                worst[key + "_seq"], worst[key + "_par"] = max(worst[key + "_seq"], e_seq.max()), max(worst[key + "_par"], e_par.max())
                assert (e_par <= e_seq + 1e-4).all(), f"step {n} {key}: time-parallel {e_par.max():.3e} vs sequential {e_seq.max():.3e} from fp64"
                assert e_par.max() <= max(1e-4, e_seq.max()) + 1e-5 or e_par.max() <= 1e-4, f"step {n} {key}: {e_par.max():.3e}"
        else:
            if n > 0:
                parity.assert_parity(sels[1][0], sels[0][0], f"step {n} selected forward variables")
                parity.assert_parity(sels[1][1], sels[0][1], f"step {n} selected prefix scores")
            parity.assert_parity(outs[1], outs[0], f"step {n} joint scores after a time-parallel selection", rtol=0)
            parity.assert_parity(procs[1].ctc_states[1], procs[0].ctc_states[1], f"step {n} log_psi")
        cand = (outs[0] + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        base = (torch.arange(B, device="cuda") * W).view(B, 1)
        ids = torch.cat([ids[(src + base).view(-1)], tok.view(-1, 1)], dim=1)
        beam_scores = top
    print(f"pscan-vs-fp64 B={B} W={W} T={T} V={V} {kind}: " + " ".join(f"{k}={v:.2e}" for k, v in worst.items()))


class _LazyBackend(parity.Backend):
    device = "cuda"

    def make_scorer(self, x_logp, lens, blank, eos, margin=0):
        from huggingface_asr_b200.decoding.ctc_scorer import CTCPrefixScoreTH

        sc = CTCPrefixScoreTH(x_logp.contiguous(), lens, blank, eos, margin)
        sc.lazy_state = True
        return sc

    def make_processor(self, logits, lens, pad, eos, margin, w, W, space=-1, trick=False, trick_w=1.0):
        from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

        return CTCRescorerLogitsProcessor(logits, lens, pad, eos, margin, w, W, space, trick, trick_w, materialize_state=False)


@pytest.mark.parametrize("name", ["steps_peaky_w3", "steps_peaky_ragged_w10", "steps_flat_w1", "steps_flat_w20", "steps_peaky_w5_v129",
                                  "steps_forced_pad", "steps_trick"])
def test_pscan_steps_vs_reference_golden(name):
    _mode(1)
    print(name, parity.replay_steps(_LazyBackend(), name))


def test_pscan_edges_decode_and_extend_vs_reference_golden():
    _mode(1)
    be = _LazyBackend()
    parity.replay_edges(be)
    parity.replay_decode(be)
    parity.replay_extend(be)


@pytest.mark.parametrize("name", [n for n in parity.PREBEAM_CASES])
def test_pscan_prebeam_vs_reference_golden(name):
    """ctcps_select_lazy_candidates (token-major posteriors, candidate ids) through the time-parallel kernel."""
    _mode(1)
    worst = parity.replay_prebeam(lambda lg, ln, w, W, S, ubi: _proc(lg, ln, W, w, pre_beam_size=S, use_beam_idx=ubi), "cuda", name,
                                  teacher_forced=name.endswith("tokens_only"))
    print(name, worst)


@pytest.mark.parametrize("pre_beam", [0, 15])
def test_pscan_native_decode_gives_the_same_1best(pre_beam):
    from huggingface_asr_b200.beam_search import joint_beam_search_native
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits

    B, W, T, V = 8, 10, 373, 5000
    logits, lens, _ = make_encoder_logits(B, T, V, "peaky", True, seed=17)
    outs = []
    for m in (0, 1):
        _mode(m)
        proc = _proc(logits.cuda(), lens.cuda(), W, **({"pre_beam_size": pre_beam} if pre_beam else {}))
        outs.append(joint_beam_search_native(proc, lambda ids, n: make_attention_scores(B * W, V, n, seed=3, scale=0.5).cuda(), B, W, V,
                                             BOS, EOS, BLANK, max_length=64, device=torch.device("cuda"), done_check_lag=0))
    assert outs[0].steps == outs[1].steps
    assert torch.equal(outs[0].sequences, outs[1].sequences), "1-best sequences differ between the sequential and the time-parallel selection"
    assert (outs[0].scores - outs[1].scores).abs().max().item() <= 1e-3
