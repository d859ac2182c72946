"""Numerics prototype (numpy, CPU) of a time-parallel lazy state selection -- groundwork for round 2, item "select scan".

`k_select_lazy_scan` recomputes, for every surviving (hypothesis, token) column, the forward recursion of the previous
step (ctc_scorer.py:148-151) with one thread walking T frames: T dependent pairs of logaddexp, ~50 us at C2, the critical
path of the pre-beam loop and of small-batch decodes (profiles/r1x_kernels_ncu.md section 5).  The recursion is an affine
map in the (logsumexp, +) semiring,

    rn' = lse(rn + xv, ph + xv)          rb' = lse(rn + bl, rb + bl)

(`ph` = phi[t-1], `xv` = x[t, tok], `bl` = x[t, blank]), i.e. state' = A (x) state (+) c with five live coefficients
(A_nn, A_bn, A_bb, c_n, c_b; rn never depends on rb), and affine maps compose associatively:

    A_nn = A2_nn + A1_nn                              c_n = lse(A2_nn + c1_n, c2_n)
    A_bn = lse(A2_bn + A1_nn, A2_bb + A1_bn)          c_b = lse(A2_bn + c1_n, A2_bb + c1_b, c2_b)
    A_bb = A2_bb + A1_bb

So a warp can own one column: lane l composes the maps of its block of F = ceil((T - start) / 32) frames (depth F), a
5-level warp scan of the 32 block maps gives every lane its incoming state, and the lane replays its block (depth F):
2F + 6 dependent logsumexp levels instead of T (30 instead of 372 at T = 373).  Everything stays in the log domain with
the reference's finite logzero (-1e10), so there is no dynamic-range problem and the "logzero class" survives as sums of
a few -1e10.

This script replays the golden traces of the reference (tests/golden/steps_*.npz: the state the reference selected for
every step) through (a) the sequential recursion in fp32, as the shipped kernel does, and (b) the block-parallel
formulation in fp32 with the same lane / block structure, and reports the error of both against the reference's fp64
run.  Test infrastructure only; nothing in the product imports it.

    python tools/pscan_prototype.py               # golden traces of the reference
    python tools/pscan_prototype.py --synthetic   # BASELINE-length streams (T = 373, 748) against an fp64 recursion
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LZ = np.float32(-1e10)
LANES = 32


def lse(a, b):
    a, b = np.float32(a), np.float32(b)
    m = np.maximum(a, b)
    return np.float32(m + np.log1p(np.exp(np.float32(-np.abs(a - b)), dtype=np.float32), dtype=np.float32))


def sequential(rn, rb, ph, xv, bl):
    out = np.empty((len(ph), 2), np.float32)
    for t in range(len(ph)):
        ls = lse(rn, rb)
        rn = np.float32(lse(rn, ph[t]) + xv[t])
        rb = np.float32(ls + bl[t])
        out[t] = rn, rb
    return out


def compose(m2, m1):
    """m2 after m1; a map is (A_nn, A_bn, A_bb, c_n, c_b)."""
    a2nn, a2bn, a2bb, c2n, c2b = m2
    a1nn, a1bn, a1bb, c1n, c1b = m1
    return (np.float32(a2nn + a1nn), lse(a2bn + a1nn, a2bb + a1bn), np.float32(a2bb + a1bb), lse(a2nn + c1n, c2n),
            lse(lse(a2bn + c1n, a2bb + c1b), c2b))


IDENT = (np.float32(0), LZ, np.float32(0), LZ, LZ)


def apply(m, rn, rb):
    ann, abn, abb, cn, cb = m
    return lse(ann + rn, cn), lse(lse(abn + rn, abb + rb), cb)


def block_parallel(rn0, rb0, ph, xv, bl):
    n = len(ph)
    F = -(-n // LANES)
    blocks = [(l * F, min(n, (l + 1) * F)) for l in range(LANES)]
    # phase 1: every lane composes the frame maps of its block
    maps = []
    for lo, hi in blocks:
        ann, abn, abb, cn, cb = IDENT
        for t in range(lo, hi):
            ann, abn, abb, cn, cb = (np.float32(xv[t] + ann), np.float32(bl[t] + lse(ann, abn)), np.float32(bl[t] + abb),
                                     np.float32(xv[t] + lse(cn, ph[t])), np.float32(bl[t] + lse(cn, cb)))
        maps.append((ann, abn, abb, cn, cb))
    # phase 2: inclusive Kogge-Stone scan over the lanes, then shift by one (exclusive)
    inc = list(maps)
    d = 1
    while d < LANES:
        inc = [compose(inc[l], inc[l - d]) if l >= d else inc[l] for l in range(LANES)]
        d *= 2
    exc = [IDENT] + inc[:-1]
    # phase 3: incoming state of every block, replay
    out = np.empty((n, 2), np.float32)
    for l, (lo, hi) in enumerate(blocks):
        if lo >= hi:
            continue
        rn, rb = apply(exc[l], rn0, rb0)
        out[lo:hi] = sequential(rn, rb, ph[lo:hi], xv[lo:hi], bl[lo:hi])
    return out


def replay(name, blank=3):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    x = g["x_padded"]
    B, T, V = x.shape
    W = int(g["W"])
    BW = B * W
    worst = {"sequential": 0.0, "block-parallel": 0.0, "seq vs par": 0.0}
    cls_bad = 0
    for n in range(int(g["n_steps"]) - 1):
        if n == 0:  # initial state (:74-85)
            r_prev = np.full((T, 2, BW), LZ, np.float32)
            for b in range(B):
                r_prev[:, 1, b * W:(b + 1) * W] = np.cumsum(x[b, :, blank], dtype=np.float32)[:, None]
        else:
            r_prev = g[f"sel_r_{n}"]
        ids_n, ids_next = g[f"input_ids_{n}"], g[f"input_ids_{n + 1}"]
        ref, ref64 = g[f"sel_r_{n + 1}"], g[f"sel_r_{n + 1}_f64"]
        ol = ids_n.shape[1] - 1
        start = max(ol, 1)
        for j in range(BW):
            b = j // W
            hs = b * W  # the processor selects with token ids only: source = first hypothesis of the utterance (:326-329)
            tok = int(ids_next[j, -1])
            last = int(ids_n[hs, -1]) == tok
            phi = r_prev[:, 1, hs] if last else lse(r_prev[:, 0, hs], r_prev[:, 1, hs])
            ph, xv, bl = phi[start - 1:T - 1], x[b, start:, tok], x[b, start:, blank]
            rn0 = x[b, 0, tok] if ol == 0 else LZ
            for kind, fn in (("sequential", sequential), ("block-parallel", block_parallel)):
                out = fn(np.float32(rn0), LZ, ph, xv, bl)
                want, want64 = ref[start:, :, j], ref64[start:, :, j]
                fin = want > -1e9
                cls_bad += int((out[~fin] > -1e9).any())
                if fin.any():
                    worst[kind] = max(worst[kind], float(np.abs(out[fin].astype(np.float64) - want64[fin]).max()))
            a, c = sequential(np.float32(rn0), LZ, ph, xv, bl), block_parallel(np.float32(rn0), LZ, ph, xv, bl)
            fin = a > -1e9
            if fin.any():
                worst["seq vs par"] = max(worst["seq vs par"], float(np.abs(a[fin] - c[fin]).max()))
    return worst, cls_bad, (B, W, T, V)


def sequential64(rn, rb, ph, xv, bl):
    out = np.empty((len(ph), 2), np.float64)
    rn, rb = np.float64(rn), np.float64(rb)
    for t in range(len(ph)):
        ls = np.logaddexp(rn, rb)
        rn = np.logaddexp(rn, np.float64(ph[t])) + np.float64(xv[t])
        rb = ls + np.float64(bl[t])
        out[t] = rn, rb
    return out


def synthetic(T, seed, columns=40):
    """BASELINE-length streams: peaky posteriors (one label or blank near 0, the rest around -12), a phi stream that decays
    like a real prefix probability, length padding (x = logzero, blank = 0) on the last frames, a leading logzero stretch."""
    rng = np.random.default_rng(seed)
    worst = {"sequential": 0.0, "block-parallel": 0.0}
    bad = 0
    for _ in range(columns):
        n = T - 1
        pad = int(rng.integers(0, T // 3))
        xv = np.where(rng.random(n) < 0.04, -rng.random(n) * 0.5, -8.0 - 8.0 * rng.random(n)).astype(np.float32)
        bl = np.where(rng.random(n) < 0.8, -rng.random(n) * 0.3, -6.0 - 4.0 * rng.random(n)).astype(np.float32)
        ph = (-0.5 * np.arange(n) * rng.random() - 3.0 * rng.random(n)).astype(np.float32)
        lead = int(rng.integers(0, 20))
        ph[:lead] = LZ
        if pad:
            xv[-pad:] = LZ
            bl[-pad:] = 0.0
        rn0 = LZ if rng.random() < 0.7 else np.float32(-rng.random() * 5)
        want = sequential64(rn0, LZ, ph, xv, bl)
        fin = want > -1e9
        for kind, fn in (("sequential", sequential), ("block-parallel", block_parallel)):
            out = fn(np.float32(rn0), LZ, ph, xv, bl)
            bad += int((out[~fin] > -1e9).any())
            if fin.any():  # tests/parity.py criterion: |new - ref| <= 1e-4 + 2e-6 |ref|; report the worst ratio to that bound
                worst[kind] = max(worst[kind], float((np.abs(out[fin] - want[fin]) / (1e-4 + 2e-6 * np.abs(want[fin]))).max()))
    return worst, bad


def main():
    if "--synthetic" in sys.argv:
        for T in (373, 748):
            worst, bad = synthetic(T, seed=T)
            print(f"synthetic T={T}: worst |err| / (1e-4 + 2e-6 |ref|) vs an fp64 recursion: sequential fp32 {worst['sequential']:.3f}, "
                  f"block-parallel fp32 {worst['block-parallel']:.3f} (must stay below 1), class errors {bad}")
        return
    names = sys.argv[1:] or ["steps_peaky_ragged_w10", "steps_flat_w20", "steps_peaky_w5_v129", "steps_forced_pad", "steps_flat_w1"]
    print(f"{'trace':28s} {'B,W,T,V':>16s}  max |err| vs the reference's fp64 run: sequential fp32 / block-parallel fp32   (seq vs par)  class errors")
    for name in names:
        worst, bad, shape = replay(name)
        print(f"{name:28s} {str(shape):>16s}  {worst['sequential']:.2e} / {worst['block-parallel']:.2e}   ({worst['seq vs par']:.2e})  {bad}")


if __name__ == "__main__":
    main()
