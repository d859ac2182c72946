#!/usr/bin/env python
"""bench.py -- utt/s of beam-10 joint CTC/attention decoding with the sm_100a CTC prefix scorer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2]

Metric (BASELINE.json): utterances/s of the joint decode + the prefix-score kernel's HBM GB/s against the
measured peak.  One "step" = one full joint CTC/attention beam-search decode of one batch of synthetic
utterances (K-a init, then per output token: state select, prefix scoring fused with the joint combine, and
the beam update of the shared harness).  Workload = BASELINE.json configs[1] (C2: 256 x 15 s utterances,
beam 10, 5000-token vocabulary) per GPU; weak scaling (each rank decodes its own batch, no collective on
the data path; the final hypotheses are gathered with NCCL inside the e2e region).

The attention decoder is model code outside the path: its log-probs come from
huggingface_asr_b200.synthetic.SyntheticDecoder, the CTC head outputs are synthetic "peaky" logits.

--impl reference times the CPU oracle port of the reference scorer (oracle/, all host threads) inside the
same harness on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from huggingface_asr_b200.beam_search import (joint_beam_search, joint_beam_search_fused, joint_beam_search_native,  # noqa: E402
                                              resolve_score_timing)
from huggingface_asr_b200.synthetic import (BLANK, BOS, CONFIGS, EOS, SyntheticDecoder, make_encoder_hidden,  # noqa: E402
                                            make_encoder_logits)

METRIC = "utt/s beam-10 joint CTC/attn decode"
UNIT = "utt/s"
MAX_LENGTH = 128  # synthetic transcripts have T//8 + 1 <= 94 tokens; the reference recipes use 512
ATT_POOL = 8


def algorithmic_bytes_per_score(B, W, T, V):
    """interface-faithful bytes of one ctcps_score launch (fp32)."""
    BW = B * W
    return (8 * T * BW * V      # write state r, both planes
            + 4 * T * B * V     # read log-posteriors once (shared by the W hyps)
            + 4 * T * B         # blank column
            + 8 * T * BW        # read r_prev
            + 4 * BW + 8 * BW   # s_prev, last ids
            + 4 * BW * V * 3)   # read attention scores, write log_psi, write joint scores


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            # nvidia-smi needs 0.1-1 s to initialise NVML, during which it holds driver locks that stall kernel launches:
            # wait for its first sample so that this happens before the warm-up, not inside a short timed region
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0 and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    from huggingface_asr_b200 import _lib
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the scorer has no CPU path (use --impl reference for the CPU oracle)")
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa_bind else None
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    B, W, T, V = args.batch or cfg.B, cfg.W, cfg.T, cfg.V
    BW = B * W
    _lib.lib()  # build / load outside the timed region

    # synthetic inputs: pinned host copy for the e2e leg, device copy for the resident leg
    logits_h, lens_h, transcripts = make_encoder_logits(B, T, V, cfg.kind, cfg.ragged, seed=20240 + 1000 * 2 + rank)
    logits_h, lens_h = logits_h.pin_memory(), lens_h.pin_memory()
    logits_d, lens_d = logits_h.to(dev), lens_h.to(dev)
    decoder = SyntheticDecoder(transcripts, W, V, MAX_LENGTH, seed=7 + rank, device=dev, pool=ATT_POOL)
    out_seq_h = torch.empty((B, MAX_LENGTH), dtype=torch.long).pin_memory()
    out_len_h = torch.empty((B,), dtype=torch.long).pin_memory()
    out_score_h = torch.empty((B,), dtype=torch.float32).pin_memory()
    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    warm = args.warmup if args.profile else max(args.warmup, 3)

    last_sequences = {}

    def measure(materialize, pre_beam=0):
        """Both legs for one state mode.  Returns a dict of raw measurements (max over ranks for the times)."""
        launches = [0]
        score_events = []
        score_ms_native = []

        def decode(lg, ln, timing=None):
            proc = CTCRescorerLogitsProcessor(lg, ln, BLANK, EOS, 0, cfg.ctc_weight, W, -1, False, 1.0, materialize_state=materialize,
                                              pre_beam_size=pre_beam)
            proc.ctc_prefix_scorer._timing = timing if (args.harness != "native" or materialize) else None
            lag = (0 if materialize else 1) if args.done_check_lag is None else args.done_check_lag
            if args.harness == "native" and not materialize:
                out = joint_beam_search_native(proc, decoder, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=dev,
                                               done_check_lag=lag, score_timing=None if timing is None else score_ms_native,
                                               fuse_topk=not args.no_fuse_topk)
            elif args.harness in ("fused", "native"):
                out = joint_beam_search_fused(proc, decoder, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=dev,
                                              done_check_lag=(0 if materialize else 1) if args.done_check_lag is None else args.done_check_lag)
            else:
                out = joint_beam_search(proc, decoder, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=dev)
            # our kernels: K-a (1) + initial state (1); per step: scoring kernel (1) [+ fused beam step (1)], plus its
            # preparation kernel (every step when materialised; first step only in lazy mode, where the select scan of
            # the previous step prepares it); all steps but the first: select (1 gather, or 2 for the lazy stage + scan)
            # beam step: the candidate kernel alone (pre-beam), or -- dense scores whose rows fit the register top-k -- the
            # per-row top-2W kernel + the candidate kernel
            two_kernel = not pre_beam and V % 4 == 0 and V <= 8192
            native = 1 if (args.harness == "native" and not materialize) else 0  # the native step also selects after the last step
            # native full-vocabulary loop: the scoring kernel ranks its tiles itself, the beam step is the list merge alone
            fused_topk = bool(native and not pre_beam and not args.no_fuse_topk and V % 4 == 0)
            beam = 0 if args.harness == "torch" else (1 if fused_topk else (2 if two_kernel else 1))
            if pre_beam:
                # K-a + transpose + initial state; per step: top-S, candidate scores, [dense scatter when the harness is not
                # sparse], [beam step]; first step: k_prep_psi; all steps but the first: select stage + scan
                dense = 0 if (args.harness != "torch" and pre_beam >= 2) else 1
                launches[0] += 3 + out.steps * (2 + dense + beam) + 1 + (out.steps - 1 + native) * 2
            else:
                launches[0] += (2 + out.steps * (1 + beam) + (out.steps if materialize else 1)
                                + (out.steps - 1 + native) * (1 if materialize else 2))
            last_sequences[(materialize, pre_beam)] = out.sequences
            return out

        # clocks are sampled from the warm-up on (same load as the timed steps), so that short timed regions still
        # get enough nvidia-smi samples; the sampler stops right after the timed region
        # rank 0 only: eight nvidia-smi pollers contend for the driver with the launch path of every rank (the pre-beam leg,
        # ~5 launches per 200 us step, dropped from 24 k to 6 k utt/s per GPU at N = 8 with one poller per rank)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        for _ in range(warm):
            decode(logits_d, lens_d)
        sync_all()
        # ---- leg 1: inputs resident in HBM, device-timed ------------------------------------------------
        launches[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        steps_total = 0
        for _ in range(args.steps):
            out = decode(logits_d, lens_d, score_events)
            steps_total += out.steps
        e1.record()
        sync_all()
        clk = clocks.stop()
        ms = e0.elapsed_time(e1)
        n_launch = launches[0]
        score_ms = [a.elapsed_time(b) for a, b in score_events] + resolve_score_timing(score_ms_native)
        res = {"ms": ms, "launches": n_launch, "score_ms": sum(score_ms) / max(len(score_ms), 1), "n_score": len(score_ms),
               "decode_steps": steps_total / args.steps, "clocks": clk, "ms_e2e": float("nan")}
        if args.profile:
            return res

        # ---- leg 2: end to end through the processor API with HOST buffers ------------------------------
        # Every step copies its encoder logits from pinned host memory (H2D) and reads its hypotheses back (D2H) inside
        # the timed region.  The H2D copy of step i+1 is enqueued on a copy stream while step i decodes (two device
        # buffers), the way a serving loop prefetches its next batch; nothing is copied outside the region.
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        bufs = [(torch.empty_like(logits_d), torch.empty_like(lens_d)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            copy_stream.wait_event(free[i % 2])  # the decode that last read this buffer is finished
            with torch.cuda.stream(copy_stream):
                bufs[i % 2][0].copy_(logits_h, non_blocking=True)
                bufs[i % 2][1].copy_(lens_h, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_run(n):
            for ev in free:
                ev.record(main)
            prefetch(0)
            for i in range(n):
                main.wait_event(ready[i % 2])
                if i + 1 < n:
                    prefetch(i + 1)
                o = decode(*bufs[i % 2])
                free[i % 2].record(main)
                if dist is not None:  # final gather of the hypotheses: the only collective of the path
                    seqs = [torch.empty_like(o.sequences) for _ in range(world)]
                    dist.all_gather(seqs, o.sequences)
                out_seq_h.copy_(o.sequences, non_blocking=True)
                out_len_h.copy_(o.lengths, non_blocking=True)
                out_score_h.copy_(o.scores, non_blocking=True)

        e2e_run(1)
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_run(args.steps)
        f1.record()
        sync_all()
        res["ms_e2e"] = f0.elapsed_time(f1)
        t = torch.tensor([res["ms"], res["ms_e2e"]], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["ms"], res["ms_e2e"] = float(t[0]), float(t[1])
        return res

    def measure_from_hidden(pre_beam=0):
        """N4 boundary: the step's inputs are the encoder's hidden states (B,T,d) in pinned host memory; the CTC head GEMM
        (split-TF32 on the tensor cores), K-a and the whole decode run inside the timed region.  Same double-buffered
        H2D prefetch as the e2e leg.  Returns (ms for args.steps steps, CUDA-event ms of one head GEMM incl. the split)."""
        from huggingface_asr_b200.ctc_head import CTCHead

        d = args.hidden_dim
        hid_h, w_h, b_h, hl_h, tr = make_encoder_hidden(B, T, V, d, cfg.kind, cfg.ragged, seed=20240 + 1000 * 2 + 500 + rank)
        hid_h, hl_h = hid_h.pin_memory(), hl_h.pin_memory()
        head = CTCHead(w_h.to(dev), b_h.to(dev))  # model weights: resident
        dec = SyntheticDecoder(tr, W, V, MAX_LENGTH, seed=11 + rank, device=dev, pool=ATT_POOL)
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        bufs = [(torch.empty(hid_h.shape, dtype=torch.float32, device=dev), torch.empty_like(lens_d)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        head_ev = []

        def prefetch(i):
            copy_stream.wait_event(free[i % 2])
            with torch.cuda.stream(copy_stream):
                bufs[i % 2][0].copy_(hid_h, non_blocking=True)
                bufs[i % 2][1].copy_(hl_h, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def run(n):
            for ev in free:
                ev.record(main)
            prefetch(0)
            for i in range(n):
                main.wait_event(ready[i % 2])
                if i + 1 < n:
                    prefetch(i + 1)
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record()
                logits = head(bufs[i % 2][0])
                h1.record()
                head_ev.append((h0, h1))
                proc = CTCRescorerLogitsProcessor(logits, bufs[i % 2][1], BLANK, EOS, 0, cfg.ctc_weight, W, -1, False, 1.0,
                                                  materialize_state=False, pre_beam_size=pre_beam)
                o = joint_beam_search_native(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=dev, done_check_lag=1)
                free[i % 2].record(main)
                if dist is not None:
                    seqs = [torch.empty_like(o.sequences) for _ in range(world)]
                    dist.all_gather(seqs, o.sequences)
                out_seq_h.copy_(o.sequences, non_blocking=True)
                out_len_h.copy_(o.lengths, non_blocking=True)
                out_score_h.copy_(o.scores, non_blocking=True)
            return o

        for _ in range(2):
            o = run(1)
        sync_all()
        ok = bool((o.lengths.cpu() == torch.tensor([len(t) - 1 for t in tr])).all())
        head_ev.clear()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        run(args.steps)
        f1.record()
        sync_all()
        t = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        head_ms = sum(a.elapsed_time(b) for a, b in head_ev) / max(len(head_ev), 1)
        return float(t[0]), head_ms, hid_h.numel() * 4 + hl_h.numel() * 8, ok

    main_mode = args.state == "materialized"
    if args.state == "pre_beam":  # diagnostic / profiling runs of the N2 path only; the bench line is always a full-vocabulary mode
        if not args.profile:
            raise SystemExit("--state pre_beam is for --profile runs; the default run reports pre-beam under its own key")
        res = measure(False, args.pre_beam)
    else:
        res = measure(main_mode)
    if args.profile:
        if rank == 0:
            emit({"profile_run": True, "state": args.state, "ms_per_step": res["ms"] / args.steps, "avg_score_ms": res["score_ms"]})
        return
    other = None if args.single_mode else measure(not main_mode)
    pre = measure(False, args.pre_beam) if (args.pre_beam > 0 and not args.single_mode) else None
    agreement = None
    if pre is not None and (False, 0) in last_sequences:
        same = (last_sequences[(False, 0)] == last_sequences[(False, args.pre_beam)]).all(dim=1).float().mean()
        agreement = float(same)

    hidden = None
    if args.hidden_dim > 0 and not args.single_mode and not args.profile:
        hidden = {"lazy": measure_from_hidden(0)}
        if pre is not None:
            hidden["pre_beam"] = measure_from_hidden(args.pre_beam)

    if rank == 0:
        peak, peak_src = peak_hbm()
        abytes = algorithmic_bytes_per_score(B, W, T, V)
        lazy_bytes = 4 * T * B * V + 4 * T * B * W + 8 * T * B * W + 12 * B * W + 12 * B * W * V  # x once + lin stream + r_prev + scores

        def roofline(r, materialized):
            true_bytes = abytes if materialized else lazy_bytes
            d = {"bound": "hbm", "kernel": "k_score_full (+ k_prep)" if materialized else "k_psi_full",
                 "achieved": true_bytes / (r["score_ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                 "traffic": load_traffic("k_score_full" if materialized else "k_psi_full"),
                 "algorithmic_bytes_per_launch": true_bytes, "avg_launch_ms": r["score_ms"], "launches_timed": r["n_score"]}
            d["frac"] = d["achieved"] / peak
            if not materialized:
                d["note"] = ("lazy state: r (T,2,BW,V) is not written; true bytes = posteriors read once + scores; "
                             "effective_* is the interface-faithful figure of SURVEY 8(d) divided by the same time")
                d["effective_achieved"] = abytes / (r["score_ms"] * 1e-3) / 1e9
                d["effective_frac"] = d["effective_achieved"] / peak
            return d

        def summary(r, materialized):
            return {"value": world * B * args.steps / (r["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms"] / args.steps,
                    "e2e": {"value": world * B * args.steps / (r["ms_e2e"] * 1e-3), "unit": UNIT,
                            "h2d_bytes_per_step": logits_h.numel() * 4 + lens_h.numel() * 8,
                            "d2h_bytes_per_step": out_seq_h.numel() * 8 + out_len_h.numel() * 8 + out_score_h.numel() * 4},
                    "gpu_launches": r["launches"], "roofline": roofline(r, materialized), "clocks": r["clocks"],
                    "decode_steps_per_utterance_batch": r["decode_steps"]}

        m = summary(res, main_mode)
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg.name}", "utterances_per_gpu": B, "beam": W, "frames": T, "vocab": V,
                       "ctc_weight": cfg.ctc_weight, "logits": cfg.kind, "state": args.state, "harness": args.harness,
                       "decode_steps_per_utterance_batch": m["decode_steps_per_utterance_batch"],
                       "attention_scores": "SyntheticDecoder: log_softmax(noise + 10*onehot(transcript[n])) (the decoder is model code outside the path)",
                       "max_length": MAX_LENGTH, "numa_bound_cpus": None if numa is None else len(numa),
                       "l2": "inputs exceed L2: posteriors are 1.9 GB and (materialized) every scorer launch writes 8*T*BW*V bytes of state"},
            "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "roofline": m["roofline"], "clocks": m["clocks"],
        }
        if other is not None:
            o = summary(other, not main_mode)
            line["lazy_state" if main_mode else "materialized_state"] = o
        if pre is not None:
            S = args.pre_beam
            line["pre_beam"] = {
                "pre_beam_size": S, "value": world * B * args.steps / (pre["ms"] * 1e-3), "unit": UNIT,
                "ms_per_step": pre["ms"] / args.steps,
                "e2e": {"value": world * B * args.steps / (pre["ms_e2e"] * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": logits_h.numel() * 4 + lens_h.numel() * 8,
                        "d2h_bytes_per_step": out_seq_h.numel() * 8 + out_len_h.numel() * 8 + out_score_h.numel() * 4},
                "gpu_launches": pre["launches"], "decode_steps_per_utterance_batch": pre["decode_steps"],
                "score_candidates_ms": pre["score_ms"], "clocks": pre["clocks"],
                "one_best_agreement_with_full_vocabulary": agreement,
                "note": ("SURVEY 8(f) N2, not the reference's behaviour: only the top-S decoder tokens of every hypothesis are "
                         "CTC-scored (ESPnet pre-beam, S = 1.5 * beam by default), states selected with hyp*V+tok; sparse fused "
                         "harness (no (BW,V) tensor); score_candidates_ms = CUDA-event time of ctcps_score_candidates"),
            }
        if hidden is not None:
            d = args.hidden_dim
            flops = 2.0 * B * T * V * d
            line["e2e_from_hidden"] = {
                k: {"value": world * B * args.steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": out_seq_h.numel() * 8 + out_len_h.numel() * 8 + out_score_h.numel() * 4,
                    "ctc_head_ms": head_ms, "ctc_head_tflops_fp32_equivalent": flops / (head_ms * 1e-3) / 1e12,
                    "transcripts_recovered": ok}
                for k, (ms, head_ms, h2d, ok) in hidden.items()}
            line["e2e_from_hidden"]["note"] = (
                f"SURVEY 8(f) N4 boundary, not the reference-facing call: host buffers hold the encoder hidden states (B,T,{d}) "
                "instead of the (B,T,V) logits; the CTC head GEMM runs on the GPU inside the timed region (operands split into "
                "TF32-exact parts, one stacked-K TF32 cuBLAS GEMM = fp32 accuracy), then K-a and the native decode loop")
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(cfg, args)
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU BEFORE it allocates pinned memory (first touch puts the
    staging buffers on that NUMA node): with several ranks per host, H2D copies from one socket's memory cap the whole job.
    Returns the CPU list, or None when NVML / sched_setaffinity is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def load_traffic(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))[f"{kernel}_dram_bytes_per_launch"]
    except Exception:
        return None


def oracle_decode(cfg, B, seed, threads=None):
    """The CPU oracle port inside the same harness on B utterances of cfg's shape.  Returns (seconds, steps)."""
    from oracle import oracle as orc

    orc.set_threads(threads or os.cpu_count() or 1)
    W, T, V = cfg.W, cfg.T, cfg.V
    logits, lens, transcripts = make_encoder_logits(B, T, V, cfg.kind, cfg.ragged, seed=seed)
    decoder = SyntheticDecoder(transcripts, W, V, MAX_LENGTH, seed=7, pool=ATT_POOL)
    t0 = time.perf_counter()
    proc = orc.OracleCTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, cfg.ctc_weight, W)
    out = joint_beam_search(proc, decoder, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH)
    return time.perf_counter() - t0, out.steps


def cpu_baseline(cfg, args):
    cores = os.cpu_count() or 1
    Bs = 2 if cores >= 16 else 1
    sec, steps = oracle_decode(cfg, Bs, seed=20240 + 2000)
    return {"value": Bs / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{Bs} utterance(s) of the {args.config} shape (T={cfg.T}, V={cfg.V}, beam {cfg.W}), full decode of {steps} steps, "
                      f"oracle/ctc_prefix_oracle.c with OpenMP on {cores} threads, {sec:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = args.batch or (2 if cores >= 16 else 1)
    for _ in range(min(args.warmup, 1)):
        oracle_decode(cfg, 1 if Bs > 1 else Bs, seed=1)
    t = 0.0
    steps = 0
    for i in range(args.steps):
        sec, st = oracle_decode(cfg, Bs, seed=20240 + 2000 + i)
        t += sec
        steps += st
    val = Bs * args.steps / t
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)),
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg.name}", "sample_utterances_per_step": Bs, "beam": cfg.W, "frames": cfg.T,
                   "vocab": cfg.V, "ctc_weight": cfg.ctc_weight, "decode_steps_per_utterance_batch": steps / args.steps},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{Bs} utterance(s) per step of the {args.config} shape, full joint decode, CPU oracle port of "
                                   f"src/decoding/ctc_scorer.py (the reference is Python/torch and cannot travel to this box)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line of the contract, on the real stdout (libraries such as NCCL print banners to fd 1, so fd 1 is
    pointed at stderr for the rest of the run)."""
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="override utterances per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--state", default="lazy", choices=["materialized", "lazy", "pre_beam"],
                    help="state mode of the headline keys; the other mode is measured too and reported under its own key")
    ap.add_argument("--single-mode", action="store_true", help="measure only --state")
    ap.add_argument("--pre-beam", type=int, default=15,
                    help="also measure pre-beam decoding with this many candidates per hypothesis (0 = skip); reported under 'pre_beam'")
    ap.add_argument("--harness", default="native", choices=["native", "fused", "torch"],
                    help="decode loop: native = one ctcps_decode_step host call per step (default; lazy state -- the materialized mode "
                         "falls back to fused), fused = processor call + one ctcps_beam_step launch per step from Python, "
                         "torch = the torch restatement of the HF loop")
    ap.add_argument("--done-check-lag", type=int, default=None,
                    help="fused harness: steps the CPU may run ahead of the GPU (default: 0 materialized, 1 lazy)")
    ap.add_argument("--hidden-dim", type=int, default=512,
                    help="also measure end to end from encoder hidden states of this width (N4 boundary; 0 = skip)")
    ap.add_argument("--no-fuse-topk", action="store_true",
                    help="native loop: write the dense joint scores and rank them in a second kernel (the round-1 step) instead of the fused per-tile top-2W")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-rank runs: do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--profile", action="store_true", help="for runs under ncu: honour a warm-up below 3 and skip the e2e/cpu legs (never a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
