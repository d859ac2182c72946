"""Per-call CUDA-event timings of the C-ABI entry points at a BASELINE config (diagnostic; bench.py is the contract)."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from huggingface_asr_b200 import _lib  # noqa: E402
from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor  # noqa: E402
from huggingface_asr_b200.synthetic import BLANK, CONFIGS, EOS, make_encoder_logits  # noqa: E402


def timeit(fn, n=10, warm=3):
    """Device time per call in ms.  The calls are captured into ONE CUDA graph and the graph is replayed, so that the host
    cost of a call (ctypes + torch.empty: 20-60 us, more than most of these kernels take) is not what gets measured; calls
    that cannot be captured fall back to eager launches (then small numbers are host-bound)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    except Exception as exc:  # noqa: BLE001
        print(f"  (graph capture failed: {type(exc).__name__}; eager timing)")
        torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--ol", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--pre-beam", type=int, default=32)
    ap.add_argument("--select-pscan", type=int, default=-1, help="0: k_select_lazy_scan (sequential), 1: k_select_lazy_pscan; default: library default")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    B, W, T, V = args.batch or cfg.B, cfg.W, cfg.T, cfg.V
    BW = B * W
    dev = torch.device("cuda")
    L = _lib.lib()
    L.ctcps_set_select_pscan(args.select_pscan)
    logits, lens, _ = make_encoder_logits(B, T, V, cfg.kind, cfg.ragged, seed=1)
    logits, lens = logits.to(dev), lens.to(dev)
    res = {}
    want = lambda k: not args.only or k in args.only.split(",")  # noqa: E731
    if want("init"):
        res["init (log-softmax+pad)"] = timeit(lambda: CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, 0.3, W, -1, False, 1.0), 5)
    att = torch.log_softmax(torch.randn(BW, V, device=dev), -1)
    ol = args.ol
    ids = torch.randint(5, V, (BW, ol + 1), device=dev)
    for mat in (True, False):
        proc = CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=mat)
        sc = proc.ctc_prefix_scorer
        # a realistic state: run step 0 then select
        proc(torch.zeros((BW, 1), dtype=torch.long, device=dev), att.clone())
        st = proc.ctc_states
        sel = sc.index_select_state(st, ids[:, -1].reshape(-1, W))
        name = "materialized" if mat else "lazy"
        if want("score"):
            res[f"score {name} (ol={ol})"] = timeit(lambda: sc._score(ids, sel, None, None, att, 0.3))
        if want("select"):
            res[f"select {name}"] = timeit(lambda: sc.index_select_state(st, ids[:, -1].reshape(-1, W)))
        del proc, sc, st, sel
        torch.cuda.empty_cache()
    if want("prebeam"):
        S = args.pre_beam
        res[f"init token-major (S={S})"] = timeit(lambda: CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, 0.3, W, -1, False, 1.0,
                                                                                     pre_beam_size=S), 5)
        proc = CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, 0.3, W, -1, False, 1.0, pre_beam_size=S)
        sc = proc.ctc_prefix_scorer
        ids0 = torch.zeros((BW, 1), dtype=torch.long, device=dev)
        cid, cj = proc.score_candidates(ids0, att.clone())
        st = proc.ctc_states
        best = (torch.arange(W, device=dev).view(1, W) * V + cid.view(B, W, S)[:, :, 0]).contiguous()  # every hyp keeps its best token
        sel = sc.index_select_state(st, best)
        ids1 = torch.cat([ids0, cid[:, :1]], 1)
        res["prebeam topk"] = timeit(lambda: proc._top_candidates(att), 20)
        cid1, ca1 = proc._top_candidates(att)
        res["prebeam score_candidates (unprepared: + k_prep_psi)"] = timeit(lambda: sc._score_candidates(ids1, sel, cid1, ca1, 0.3), 20)
        res["prebeam select (stage + scan)"] = timeit(lambda: sc.index_select_state(st, best), 20)
        res["prebeam dense __call__ (topk + score + to_dense)"] = timeit(lambda: (setattr(proc, "ctc_states", None), proc(ids0, att))[1], 20)
        maxlen = 128
        idc = torch.randint(5, V, (BW, maxlen), device=dev)
        idn = torch.empty_like(idc)
        bs = torch.randn(B, W, device=dev)
        ps = torch.full((B, W), float("-inf"), device=dev)
        pl = torch.zeros(B, W, dtype=torch.long, device=dev)
        pq = torch.zeros(B, W, maxlen, dtype=torch.long, device=dev)
        done = torch.zeros(B, dtype=torch.uint8, device=dev)
        n = ctypes.c_size_t(0)
        L.ctcps_beam_step_workspace_bytes(B, W, ctypes.byref(n))
        ws = torch.zeros((n.value + 15) // 16 * 2, dtype=torch.int64, device=dev)
        bo = torch.empty((B, W), dtype=torch.long, device=dev)
        st_ = torch.cuda.current_stream().cuda_stream
        res["prebeam beam_step_candidates"] = timeit(lambda: _lib.check(L.ctcps_beam_step_candidates(
            cj.data_ptr(), cid.data_ptr(), S, bs.data_ptr(), idc.data_ptr(), idn.data_ptr(), maxlen, ol + 1, B, W, V, EOS, BLANK,
            float(ol + 1), ps.data_ptr(), pl.data_ptr(), pq.data_ptr(), maxlen, done.data_ptr(), ws.data_ptr(), ws.numel() * 8, None, 0, 0,
            bo.data_ptr(), torch.cuda.current_stream().cuda_stream), "beam_cand"), 20)
        del proc, sc, st, sel
        torch.cuda.empty_cache()
    if want("beam"):
        joint = torch.randn(BW, V, device=dev)
        maxlen = 128
        idc = torch.randint(5, V, (BW, maxlen), device=dev)
        idn = torch.empty_like(idc)
        bs = torch.randn(B, W, device=dev)
        ps = torch.full((B, W), float("-inf"), device=dev)
        pl = torch.zeros(B, W, dtype=torch.long, device=dev)
        pq = torch.zeros(B, W, maxlen, dtype=torch.long, device=dev)
        done = torch.zeros(B, dtype=torch.uint8, device=dev)
        n = ctypes.c_size_t(0)
        L.ctcps_beam_step_workspace_bytes(B, W, ctypes.byref(n))
        ws = torch.zeros((n.value + 15) // 16 * 2, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        res["beam_step"] = timeit(lambda: _lib.check(L.ctcps_beam_step(joint.data_ptr(), bs.data_ptr(), idc.data_ptr(), idn.data_ptr(), maxlen,
                                                                       ol + 1, B, W, V, EOS, BLANK, float(ol + 1), ps.data_ptr(), pl.data_ptr(),
                                                                       pq.data_ptr(), maxlen, done.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                                                       None, 0, 0, None, torch.cuda.current_stream().cuda_stream), "beam"), 20)
        res["torch topk(2W) of (B, W*V) for comparison"] = timeit(lambda: joint.view(B, W * V).topk(2 * W, dim=1), 20)
    for k, v in res.items():
        print(f"{k:48s} {v * 1e3:10.1f} us")


if __name__ == "__main__":
    main()
