"""Device timeline of one native decode (CUPTI through torch.profiler: every kernel of the process, ours included).

nsys is not in the image; ncu serialises launches and so cannot show what overlaps or how long the device idles between
two kernels.  This tool runs `--decodes` warm decodes of a BASELINE config under the profiler and writes

    gpurun_out/<tag>_timeline.csv      name, stream, start_us, dur_us   (every kernel / memcpy of the profiled decodes)
    gpurun_out/<tag>_timeline.txt      per-step schedule of the steady state: for each kernel of a middle decode step its
                                       offset from the step's scoring kernel, duration and stream; idle time of the main
                                       chain; per-decode fixed costs (K-a, set-up, finalisation)

Never a bench value (the profiler adds launch overhead); it explains the difference between ms_per_step and the sum of
the kernel durations.
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--state", default="lazy", choices=["lazy", "pre_beam", "materialized"])
    ap.add_argument("--decodes", type=int, default=2)
    ap.add_argument("--tag", default="tl")
    ap.add_argument("--no-fuse-topk", action="store_true")
    a = ap.parse_args()
    import bench

    args = bench.main.__globals__["argparse"].Namespace(
        gpus=1, steps=1, warmup=3, impl="ours", config=a.config, batch=None, no_cpu_baseline=True, state=a.state, single_mode=True,
        pre_beam=15, harness="native", done_check_lag=None, hidden_dim=0, extra_configs="", c5_utterances=0, no_drop_in=True,
        no_fuse_topk=a.no_fuse_topk, no_numa_bind=True, profile=True)
    R = bench.Runner(args)
    wl = bench.Workload(a.config, 0, R.dev)
    mat = a.state == "materialized"
    pre = 15 if a.state == "pre_beam" else 0

    def one():
        return R.decode(wl, wl.logits_d, wl.lens_d, mat, pre, "native")

    for _ in range(3):
        out = one()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.decodes):
            out = one()
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            evs.append((e.name, int(getattr(e, "device_index", 0)), e.time_range.start, e.time_range.end - e.time_range.start))
    # stream ids are only in the chrome trace: take them from there
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    trace = os.path.join(ROOT, "gpurun_out", f"{a.tag}_trace.json")
    prof.export_chrome_trace(trace)
    import json

    tr = json.load(open(trace))
    ks = []
    for e in tr["traceEvents"]:
        if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e:
            ks.append((e["name"], e.get("args", {}).get("stream", -1), float(e["ts"]), float(e["dur"])))
    os.remove(trace)
    ks.sort(key=lambda r: r[2])
    t0 = ks[0][2]
    with open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_timeline.csv"), "w") as f:
        f.write("name,stream,start_us,dur_us\n")
        for n, s, ts, d in ks:
            f.write(f"\"{n[:90]}\",{s},{ts - t0:.3f},{d:.3f}\n")

    # ---- per-step schedule ------------------------------------------------------------------------------------------------
    lines = []
    anchor = "k_psi_full" if a.state == "lazy" else ("k_cand" if a.state == "pre_beam" else "k_score_full")
    idx = [i for i, r in enumerate(ks) if anchor in r[0]]
    lines.append(f"{a.config} {a.state}: {len(ks)} device activities, {len(idx)} '{anchor}' launches in {a.decodes} decode(s) of {out.steps} steps")
    if len(idx) > 8:
        starts = [ks[i][2] for i in idx]
        periods = [b - c for b, c in zip(starts[1:], starts[:-1])]
        per = sorted(periods)
        lines.append(f"step period (anchor start to anchor start): median {per[len(per) // 2]:.1f} us, min {per[0]:.1f}, max {per[-1]:.1f}")
        mid = idx[len(idx) // 4]
        nxt = idx[len(idx) // 4 + 1]
        lines.append(f"one steady-state step (activities from one anchor start to the next; offsets in us):")
        base = ks[mid][2]
        busy_main = 0.0
        main_stream = ks[mid][1]
        prev_end = None
        for n, s, ts, d in ks[mid:nxt]:
            gap = "" if prev_end is None or s != main_stream else f"  idle before: {ts - prev_end:6.1f}"
            lines.append(f"  +{ts - base:8.1f}  {d:8.1f} us  stream {s:<4} {n[:70]}{gap}")
            if s == main_stream:
                busy_main += d
                prev_end = ts + d
        period = ks[nxt][2] - base
        lines.append(f"  period {period:.1f} us, main stream busy {busy_main:.1f} us, idle {period - busy_main:.1f} us")
        # per-decode fixed cost: first anchor of a decode minus the end of the previous decode's last anchor
        big = [i for i, p in enumerate(periods) if p > 3 * per[len(per) // 2]]
        for i in big:
            lines.append(f"between decodes: {periods[i]:.1f} us from the last step's anchor start to the next decode's first")
        tot = ks[-1][2] + ks[-1][3] - ks[0][2]
        lines.append(f"whole profiled region {tot / 1e3:.3f} ms = {tot / a.decodes / 1e3:.3f} ms per decode")
    txt = "\n".join(lines)
    print(txt)
    open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_timeline.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
