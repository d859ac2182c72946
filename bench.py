#!/usr/bin/env python
"""bench.py -- utt/s of beam-10 joint CTC/attention decoding with the sm_100a CTC prefix scorer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2]

Metric (BASELINE.json): utterances/s of the joint decode + the prefix-score kernel's HBM GB/s against the
measured peak.  One "step" = one full joint CTC/attention beam-search decode of one batch of synthetic
utterances (K-a init, then per output token: state select, prefix scoring fused with the joint combine, and
the beam update).  Headline workload = BASELINE.json configs[1] (C2: 256 x 15 s utterances, beam 10,
5000-token vocabulary) per GPU; weak scaling (each rank decodes its own batch, no collective on the data
path; the final hypotheses are gathered with NCCL inside the e2e region).

Besides the headline keys the line carries (all measured in the same run, every one with its own config):
  lazy_state / materialized_state   the other state mode of the scorer on the headline workload
  pre_beam                          ESPnet pre-beam decoding
  e2e_from_hidden                   host buffers hold encoder hidden states, CTC head on the GPU (N4 boundary)
  configs                           C1 / C3 / C4 of BASELINE.json, short runs, each with its own roofline
  drop_in                           the reference-facing call -- CTCRescorerLogitsProcessor.__call__ -- under the torch
                                    restatement of the HF loop, under the same loop with the beam update in one kernel, and
                                    under transformers' own generate()
  c5_job                            BASELINE.json configs[4]: 8192 ragged utterances sharded over the ranks in job form
                                    (shard_utterances -> decode_shard -> gather_hypotheses), strong scaling
  cpu_baseline                      the CPU oracle port of the reference scorer on the host cores (N = 1 only)

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
MAX_LENGTH = 128  # synthetic transcripts have T//8 + 1 <= 94 tokens; the reference recipes use 512 (only sizes the id buffers)
ATT_POOL = 8
READ_STREAM_GBS = 7250.0  # best read-only stream measured on this pool's B200s (profiles/r1_membw_read_patterns.log)


def algorithmic_bytes_per_score(B, W, T, V):
    """interface-faithful bytes of one ctcps_score launch (fp32)."""
    BW = B * W
    return (8 * T * BW * V      # write state r, both planes
            + 4 * T * B * V     # read log-posteriors once (shared by the W hyps)
            + 4 * T * B         # blank column
            + 8 * T * BW        # read r_prev
            + 4 * BW + 8 * BW   # s_prev, last ids
            + 4 * BW * V * 3)   # read attention scores, write log_psi, write joint scores


def lazy_bytes_per_score(B, W, T, V, fused_topk):
    """True bytes of one lazy scoring launch: posteriors once + lin stream + attention scores (+ the dense outputs)."""
    BW = B * W
    b = 4 * T * B * V + 4 * T * BW + 12 * BW + 4 * BW * V
    return b + (0 if fused_topk else 8 * BW * V)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def peak_tensor():
    """Dense bf16 tensor peak of this pool's B200s (burst figure: the kernel is timed alone), else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["bf16_tflops"]), "measured (MEASURED_PEAKS.json, cuBLAS bf16 burst)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md)"


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


class Workload:
    """Synthetic inputs of one BASELINE config on one rank: pinned host copy (e2e legs), device copy (resident legs), decoder."""

    def __init__(self, name, rank, dev, batch=None, ragged=None):
        self.name, self.cfg = name, CONFIGS[name]
        cfg = self.cfg
        self.B, self.W, self.T, self.V = batch or cfg.B, cfg.W, cfg.T, cfg.V
        self.ragged = cfg.ragged if ragged is None else ragged
        self.dev = dev
        idx = sorted(CONFIGS).index(name)
        logits_h, lens_h, self.transcripts = make_encoder_logits(self.B, self.T, self.V, cfg.kind, self.ragged, seed=20240 + 1000 * (idx + 1) + rank)
        self.logits_h, self.lens_h = logits_h.pin_memory(), lens_h.pin_memory()
        self.logits_d, self.lens_d = self.logits_h.to(dev), self.lens_h.to(dev)
        self.decoder = SyntheticDecoder(self.transcripts, self.W, self.V, MAX_LENGTH, seed=7 + rank, device=dev, pool=ATT_POOL)
        self.out_seq_h = torch.empty((self.B, MAX_LENGTH), dtype=torch.long).pin_memory()
        self.out_len_h = torch.empty((self.B,), dtype=torch.long).pin_memory()
        self.out_score_h = torch.empty((self.B,), dtype=torch.float32).pin_memory()

    @property
    def h2d_bytes(self):
        return self.logits_h.numel() * 4 + self.lens_h.numel() * 8

    @property
    def d2h_bytes(self):
        return self.out_seq_h.numel() * 8 + self.out_len_h.numel() * 8 + self.out_score_h.numel() * 4

    def transcripts_recovered(self, out):
        want = torch.tensor([len(t) - 1 for t in self.transcripts])
        return bool((out.lengths.cpu() == want).all())


def count_launches(steps, harness, materialize, pre_beam, V, fused_topk):
    """Our kernels launched by one decode of `steps` decoder steps (bench.py's gpu_launches claim; the ncu launch lists under
    profiles/ are the evidence).  K-a (1) + initial state (1); per step: scoring kernel (1) + beam step (native / fused
    harness: the list merge alone with the fused top-2W, else per-row top-2W + candidate kernel); the preparation kernel every
    step when materialised, on the first step only in lazy mode (the select of the previous step prepares the next call);
    all steps but the first: select (1 gather, or 3 for the lazy stage + scan + lin range; the range kernel also follows the
    first step's preparation); the native loop also selects after the last step."""
    two_kernel = not pre_beam and V % 4 == 0 and V <= 8192
    native = 1 if (harness == "native" and not materialize) else 0
    beam = 0 if harness in ("torch", "hf") else (1 if fused_topk else (2 if two_kernel else 1))
    if pre_beam:
        # K-a + transpose + initial state; per step: top-S, candidate scores, [dense scatter when the harness is not sparse],
        # [beam step]; first step: k_prep_psi; all steps but the first: select stage + scan
        dense = 0 if (harness not in ("torch", "hf") and pre_beam >= 2) else 1
        return 3 + steps * (2 + dense + beam) + 2 + (steps - 1 + native) * 3
    return 2 + steps * (1 + beam) + (steps if materialize else 2) + (steps - 1 + native) * (1 if materialize else 3)


class Runner:
    def __init__(self, args):
        from huggingface_asr_b200 import _lib

        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the scorer has no CPU path (use --impl reference for the CPU oracle)")
        self.args = args
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        self.dev = torch.device(f"cuda:{self.local}")
        torch.cuda.set_device(self.dev)
        # the staging buffers of the e2e legs should sit on the GPU's own NUMA node for every world size (first touch); the
        # original affinity is restored before the CPU baseline, which uses every host core
        self.affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        self.numa = bind_to_gpu_numa_node(self.local) if not args.no_numa_bind else None
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        _lib.lib()  # build / load outside the timed region
        self.skip_done = bool(_lib.lib().ctcps_set_skip_done(-1))
        self.frame_window = bool(_lib.lib().ctcps_set_frame_window(-1))
        self.peak, self.peak_src = peak_hbm()
        self.last_sequences = {}

    def sync_all(self):
        torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, *vals):
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def fused_topk(self, wl, harness, materialize, pre_beam):
        return bool(harness == "native" and not materialize and not pre_beam and not self.args.no_fuse_topk and wl.V % 4 == 0)

    # ------------------------------------------------------------------------------------------------
    def decode(self, wl, lg, ln, materialize, pre_beam, harness, timing=None, native_timing=None):
        from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

        args = self.args
        B, W, V = wl.B, wl.W, wl.V
        proc = CTCRescorerLogitsProcessor(lg, ln, BLANK, EOS, 0, wl.cfg.ctc_weight, W, -1, False, 1.0, materialize_state=materialize,
                                          pre_beam_size=pre_beam)
        proc.ctc_prefix_scorer._timing = timing if (harness != "native" or materialize) else None
        lag = (0 if materialize else 1) if args.done_check_lag is None else args.done_check_lag
        if harness == "native" and not materialize:
            return joint_beam_search_native(proc, wl.decoder, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=self.dev,
                                            done_check_lag=lag, score_timing=native_timing, fuse_topk=not args.no_fuse_topk)
        if harness in ("fused", "native"):
            return joint_beam_search_fused(proc, wl.decoder, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=self.dev,
                                           done_check_lag=lag)
        return joint_beam_search(proc, wl.decoder, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=self.dev)

    def measure(self, wl, materialize, pre_beam=0, harness=None, steps=None, warm=None, e2e=True, clocks=True, decode_fn=None):
        """Resident leg (inputs in HBM, device-timed) and end-to-end leg (host buffers, H2D + D2H inside the timed region) of one
        mode of one workload.  Times are the max over ranks.  decode_fn(lg, ln, timing) overrides the decode."""
        args = self.args
        harness = harness or args.harness
        steps = steps or args.steps
        warm = (args.warmup if args.profile else max(args.warmup, 3)) if warm is None else warm
        launches = [0]
        score_events, native_events = [], []
        fused_tk = self.fused_topk(wl, harness, materialize, pre_beam)
        from huggingface_asr_b200 import _lib as _libmod

        L_ = _libmod.lib()

        def run(lg, ln, timing=None):
            if decode_fn is not None:
                out = decode_fn(lg, ln, timing)
            else:
                out = self.decode(wl, lg, ln, materialize, pre_beam, harness, timing, None if timing is None else native_events)
            launches[0] += count_launches(out.steps, harness, materialize, pre_beam, wl.V, fused_tk)
            self.last_sequences[(wl.name, materialize, pre_beam, harness)] = out.sequences
            return out

        # clocks are sampled from the warm-up on (same load as the timed steps), so that short timed regions still get enough
        # nvidia-smi samples; rank 0 only: eight pollers contend for the driver with the launch path of every rank
        sampler = ClockSampler(self.local)
        if self.rank == 0 and clocks:
            sampler.start()
        for _ in range(warm):
            run(wl.logits_d, wl.lens_d)
        self.sync_all()
        launches[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the lazy scoring kernel reports how many chunks it actually streamed (finished utterances, padded frames and frames
        # whose weights underflowed to zero are left out): the roofline divides THOSE bytes by the time of the calls
        stream_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
        L_.ctcps_set_stream_counter(stream_counter.data_ptr())
        try:
            self.sync_all()
            e0.record()
            steps_total = 0
            out = None
            for _ in range(steps):
                out = run(wl.logits_d, wl.lens_d, score_events)
                steps_total += out.steps
            e1.record()
            self.sync_all()
        finally:  # never leave the library pointing at a tensor that is about to be freed
            torch.cuda.synchronize(self.dev)
            L_.ctcps_set_stream_counter(None)
        streamed_chunks = int(stream_counter.item())
        clk = sampler.stop() if (self.rank == 0 and clocks) else None
        ms = e0.elapsed_time(e1)
        score_ms = [a.elapsed_time(b) for a, b in score_events] + resolve_score_timing(native_events)
        score_ms_all = list(score_ms)
        # The native loop does not score finished utterances: a call on a batch that is (partly) finished streams less than the
        # algorithmic bytes the roofline divides by.  All utterances of a bench batch finish on the same step, so such calls are
        # the last one or two of a decode: the roofline averages the FULL calls only (>= half the median) and counts the others.
        n_all = len(score_ms)
        if n_all:
            med = sorted(score_ms)[n_all // 2]
            score_ms = [t for t in score_ms if t >= 0.5 * med]
        res = {"ms": ms, "launches": launches[0], "score_ms": sum(score_ms) / max(len(score_ms), 1), "n_score": len(score_ms),
               "n_score_skipped": n_all - len(score_ms), "score_ms_total": sum(score_ms_all), "n_score_all": n_all,
               "streamed_chunks": streamed_chunks, "chunk_bytes": int(L_.ctcps_stream_chunk_bytes(wl.W)),
               "decode_steps": steps_total / steps, "clocks": clk, "ms_e2e": float("nan"), "steps": steps, "warm": warm,
               "fused_topk": fused_tk, "harness": harness, "transcripts_recovered": wl.transcripts_recovered(out)}
        if args.profile or not e2e:
            res["ms"] = self.max_over_ranks(res["ms"])[0]
            return res

        # ---- end to end through the API with HOST buffers ------------------------------------------------
        # Every step copies its encoder logits from pinned host memory (H2D) and reads its hypotheses back (D2H) inside the
        # timed region.  The H2D copy of step i+1 is enqueued on a copy stream while step i decodes (two device buffers), the
        # way a serving loop prefetches its next batch; nothing is copied outside the region.
        dev, dist, world = self.dev, self.dist, self.world
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        bufs = [(torch.empty_like(wl.logits_d), torch.empty_like(wl.lens_d)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            copy_stream.wait_event(free[i % 2])  # the decode that last read this buffer is finished
            with torch.cuda.stream(copy_stream):
                bufs[i % 2][0].copy_(wl.logits_h, non_blocking=True)
                bufs[i % 2][1].copy_(wl.lens_h, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_run(n):
            for ev in free:
                ev.record(main)
            prefetch(0)
            for i in range(n):
                main.wait_event(ready[i % 2])
                if i + 1 < n:
                    prefetch(i + 1)
                o = run(*bufs[i % 2])
                free[i % 2].record(main)
                if dist is not None:  # final gather of the hypotheses: the only collective of the path
                    seqs = [torch.empty_like(o.sequences) for _ in range(world)]
                    dist.all_gather(seqs, o.sequences)
                wl.out_seq_h.copy_(o.sequences, non_blocking=True)
                wl.out_len_h.copy_(o.lengths, non_blocking=True)
                wl.out_score_h.copy_(o.scores, non_blocking=True)

        e2e_run(1)
        self.sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_run(steps)
        f1.record()
        self.sync_all()
        res["ms_e2e"] = f0.elapsed_time(f1)
        res["ms"], res["ms_e2e"] = self.max_over_ranks(res["ms"], res["ms_e2e"])
        del bufs
        return res

    def measure_from_hidden(self, wl, pre_beam=0):
        """N4 boundary: the step's inputs are the encoder's hidden states (B,T,d) in pinned host memory; the CTC head (GEMM on
        the tensor cores + log-softmax + padding) and the whole decode run inside the timed region.  Same double-buffered
        H2D prefetch as the e2e leg.  Returns (ms for args.steps steps, CUDA-event ms of the head, h2d bytes, ok, implementation)."""
        from huggingface_asr_b200.ctc_head import CTCHead
        from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

        args, dev, dist, world = self.args, self.dev, self.dist, self.world
        B, W, T, V = wl.B, wl.W, wl.T, wl.V
        d = args.hidden_dim
        hid_h, w_h, b_h, hl_h, tr = make_encoder_hidden(B, T, V, d, wl.cfg.kind, wl.ragged, seed=20240 + 1000 * 2 + 500 + self.rank)
        hid_h, hl_h = hid_h.pin_memory(), hl_h.pin_memory()
        head = CTCHead(w_h.to(dev), b_h.to(dev))  # model weights: resident
        dec = SyntheticDecoder(tr, W, V, MAX_LENGTH, seed=11 + self.rank, device=dev, pool=ATT_POOL)
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        bufs = [(torch.empty(hid_h.shape, dtype=torch.float32, device=dev), torch.empty_like(wl.lens_d)) for _ in range(2)]
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
            o = None
            for i in range(n):
                main.wait_event(ready[i % 2])
                if i + 1 < n:
                    prefetch(i + 1)
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record()
                proc = CTCRescorerLogitsProcessor.from_encoder_hidden_states(bufs[i % 2][0], head, bufs[i % 2][1], BLANK, EOS, 0,
                                                                             wl.cfg.ctc_weight, W, -1, False, 1.0,
                                                                             materialize_state=False, pre_beam_size=pre_beam)
                h1.record()
                head_ev.append((h0, h1))
                o = joint_beam_search_native(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=dev, done_check_lag=1,
                                             fuse_topk=not args.no_fuse_topk)
                free[i % 2].record(main)
                if dist is not None:
                    seqs = [torch.empty_like(o.sequences) for _ in range(world)]
                    dist.all_gather(seqs, o.sequences)
                wl.out_seq_h.copy_(o.sequences, non_blocking=True)
                wl.out_len_h.copy_(o.lengths, non_blocking=True)
                wl.out_score_h.copy_(o.scores, non_blocking=True)
            return o

        for _ in range(3):
            o = run(1)
        self.sync_all()
        ok = bool((o.lengths.cpu() == torch.tensor([len(t) - 1 for t in tr])).all())
        head_ev.clear()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        run(args.steps)
        f1.record()
        self.sync_all()
        ms = self.max_over_ranks(f0.elapsed_time(f1))[0]
        head_ms = sum(a.elapsed_time(b) for a, b in head_ev) / max(len(head_ev), 1)
        return ms, head_ms, hid_h.numel() * 4 + hl_h.numel() * 8, ok, getattr(head, "implementation", "cublas")

    def head_roofline(self, wl):
        """Tensor-bound roofline of the CTC head's GEMM (N4): CUDA events around `head(hidden)` -- operand split + k_head_gemm
        writing raw logits, no normalisation pass -- on resident hidden states; flops = the three fp16 products the kernel issues."""
        from huggingface_asr_b200.ctc_head import CTCHead

        args, dev = self.args, self.dev
        B, T, V, d = wl.B, wl.T, wl.V, args.hidden_dim
        hid_h, w_h, b_h, _, _ = make_encoder_hidden(B, T, V, d, wl.cfg.kind, wl.ragged, seed=20240 + 1000 * 2 + 500 + self.rank)
        head = CTCHead(w_h.to(dev), b_h.to(dev))
        if getattr(head, "implementation", "cublas") != "tcgen05":
            return None
        hid = hid_h.to(dev)
        for _ in range(3):
            z = head(hid)
        del z
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            z = head(hid)
        e1.record()
        torch.cuda.synchronize(dev)
        del z, hid
        ms = e0.elapsed_time(e1) / n
        peak, src = peak_tensor()
        flops = 3 * 2.0 * B * T * V * d
        return {"bound": "tensor", "kernel": "k_split_blocked + k_head_gemm (3xFP16 tcgen05 UMMA, raw logits out)", "achieved": flops / (ms * 1e-3) / 1e12,
                "peak": peak, "unit": "TFLOP/s", "peak_source": src, "frac": flops / (ms * 1e-3) / 1e12 / peak, "avg_launch_ms": ms,
                "flops_per_launch": flops, "traffic": None,
                "note": "flops = 3 products x 2 B T V d of fp16 tensor work for fp32-grade logits (an fp32 GEMM of the same shape is a third of it)"}

    # ------------------------------------------------------------------------------------------------
    def roofline(self, wl, r, materialized):
        abytes = algorithmic_bytes_per_score(wl.B, wl.W, wl.T, wl.V)
        true_bytes = abytes if materialized else lazy_bytes_per_score(wl.B, wl.W, wl.T, wl.V, r["fused_topk"])
        kernel = "k_score_full (+ k_prep)" if materialized else ("k_psi_full<TOPK> (fused per-tile top-2W)" if r["fused_topk"] else "k_psi_full")
        d = {"bound": "hbm", "kernel": kernel, "achieved": true_bytes / (r["score_ms"] * 1e-3) / 1e9, "peak": self.peak, "unit": "GB/s",
             "peak_source": self.peak_src,
             "traffic": load_traffic(("k_score_full" if materialized else "k_psi_full") + ("" if wl.name == "C2" else "_" + wl.name)),
             "algorithmic_bytes_per_launch": true_bytes, "avg_launch_ms": r["score_ms"], "launches_timed": r["n_score"],
             "launches_on_finished_batches_not_counted": r.get("n_score_skipped", 0)}
        if not materialized and r.get("streamed_chunks", 0) > 0 and r.get("n_score_all", 0) > 0:
            # bytes the launches of the timed region really moved: the chunks the kernel counted (8 frames x 512 tokens of
            # posteriors + the group's weights each) + per launch the decoder scores it reads and the small per-hypothesis vectors
            BW = wl.B * wl.W
            fixed = 12 * BW + 4 * BW * wl.V + (0 if r["fused_topk"] else 8 * BW * wl.V)
            moved = r["streamed_chunks"] * r["chunk_bytes"] + r["n_score_all"] * fixed
            d["achieved"] = moved / (r["score_ms_total"] * 1e-3) / 1e9
            d["avg_launch_ms"] = r["score_ms_total"] / r["n_score_all"]
            d["launches_timed"] = r["n_score_all"]
            d["launches_on_finished_batches_not_counted"] = 0
            d["traffic"] = None  # per-launch DRAM bytes vary with the window; one captured launch: profiles/r3h_psi_ncu.md (596 MB)
            d["streamed_bytes_per_launch"] = moved / r["n_score_all"]
            d["streamed_fraction_of_all_frames"] = moved / (r["n_score_all"] * true_bytes)
            d["algorithmic_bytes_per_launch_note"] = ("all frames from the prefix length to T; the kernel leaves out the chunks whose weights "
                                                      "exp(r_sum - offset) are exactly zero, the padded frames and finished utterances "
                                                      "(bit-identical scores): achieved = streamed bytes (counted by the kernel) / time of all calls")
        d["frac"] = d["achieved"] / self.peak
        d["frac_of_read_stream"] = d["achieved"] / READ_STREAM_GBS
        d["read_stream_gbs"] = READ_STREAM_GBS
        if not materialized:
            d["note"] = ("lazy state: r (T,2,BW,V) is not written; true bytes = posteriors read once + scores; effective_* is the "
                         "interface-faithful figure of SURVEY 8(d) divided by the same time; read_stream_gbs = best read-only stream "
                         "measured on this pool (the copy peak counts read + write)")
            d["effective_achieved"] = abytes / (d["avg_launch_ms"] * 1e-3) / 1e9
            d["effective_frac"] = d["effective_achieved"] / self.peak
        return d

    def summary(self, wl, r, materialized, with_e2e=True):
        n = self.world * wl.B * r["steps"]
        s = {"value": n / (r["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms"] / r["steps"], "steps": r["steps"], "warmup": r["warm"],
             "gpu_launches": r["launches"], "roofline": self.roofline(wl, r, materialized),
             "decode_steps_per_utterance_batch": r["decode_steps"], "transcripts_recovered": r["transcripts_recovered"],
             "harness": r["harness"], "fused_topk": r["fused_topk"]}
        if r["clocks"] is not None:
            s["clocks"] = r["clocks"]
        if with_e2e and r["ms_e2e"] == r["ms_e2e"]:
            s["e2e"] = {"value": n / (r["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes}
        return s

    def workload_config(self, wl, r, state):
        return {"workload": f"{wl.name}: {wl.cfg.name}", "utterances_per_gpu": wl.B, "beam": wl.W, "frames": wl.T, "vocab": wl.V,
                "ctc_weight": wl.cfg.ctc_weight, "logits": wl.cfg.kind, "lengths": "ragged (0.6 T .. T)" if wl.ragged else "equal (T)",
                "state": state, "harness": r["harness"], "fused_topk": r["fused_topk"],
                "finished_utterances": ("not scored (native loop, ctcps_set_skip_done; the returned hypotheses are bit-identical)"
                                        if r["harness"] == "native" and r["fused_topk"] and self.skip_done else
                                        "scored until the whole batch is finished, like HF's loop"),
                "frames_streamed": ("only the 8-frame chunks in which some hypothesis' weight exp(r_sum - offset) is nonzero in fp32 "
                                    "(ctcps_set_frame_window; exact: the others add 0 to every score)" if self.frame_window and state == "lazy"
                                    else "every frame from the prefix length to T"),
                "decode_steps_per_utterance_batch": r["decode_steps"]}

    # ------------------------------------------------------------------------------------------------
    def drop_in(self, wl):
        """The reference-facing call on the record: CTCRescorerLogitsProcessor.__call__ once per output token
        (a) under beam_search.joint_beam_search, the torch restatement of the HF loop's contract with the processor, and
        (b) under transformers' own generate() through JointCTCAttentionGenerationMixin (reference
        ctc_encoder_plus_autoregressive_decoder.py:360-404,450-482).  Both resident and end to end (host logits)."""
        out = {}
        steps = max(2, min(self.args.steps, 3))
        r = self.measure(wl, False, 0, harness="torch", steps=steps, warm=2, clocks=False)
        out["torch_harness"] = self.summary(wl, r, False)
        out["torch_harness"]["note"] = "beam_search.joint_beam_search: processor(input_ids, log_probs) + torch top-2W / gathers per step"
        r = self.measure(wl, False, 0, harness="fused", steps=steps, warm=2, clocks=False)
        out["fused_harness"] = self.summary(wl, r, False)
        out["fused_harness"]["note"] = ("beam_search.joint_beam_search_fused: the same processor(input_ids, log_probs) call per step, the beam "
                                        "update between two calls is one ctcps_beam_step launch, state selection prefetched on a side stream")
        try:
            hf = HFGenerate(wl, self)
            r = self.measure(wl, False, 0, harness="hf", steps=steps, warm=2, clocks=False, decode_fn=hf)
            out["hf_generate"] = self.summary(wl, r, False)
            out["hf_generate"]["note"] = ("transformers.generate() (beam search, KV-cache interface) around a stub decoder that returns the "
                                          "SyntheticDecoder log-probs; the processor is built per call by the mixin's _get_logits_processor")
        except Exception as exc:  # noqa: BLE001  -- a transformers API drift must not take the bench line down
            out["hf_generate"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        return out

    def c5_job(self):
        """BASELINE.json configs[4] in job form: 8192 ragged utterances (32 copies of a pool of 256 distinct ones), sharded over the
        ranks by sharding.shard_utterances (longest first, round robin), decoded batch by batch with the native loop, hypotheses
        gathered with one all_gather at the end.  Strong scaling: the job is the same at every N."""
        from huggingface_asr_b200 import sharding
        from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

        args, dev, rank, world = self.args, self.dev, self.rank, self.world
        cfg = CONFIGS["C5"]
        W, T, V = cfg.W, cfg.T, cfg.V
        N, P, Bb = args.c5_utterances, 256, cfg.B
        # the same pool on every rank (seed without the rank): utterance i of the job is pool utterance i % P
        pool_logits, pool_lens, pool_tr = make_encoder_logits(P, T, V, cfg.kind, True, seed=20240 + 5000)
        pool_d = pool_logits.to(dev)
        del pool_logits
        lengths = [int(pool_lens[i % P]) for i in range(N)]
        shards = sharding.shard_utterances(lengths, world)
        mine = shards[rank]
        # decoder noise that does not depend on the batch row: an utterance decodes the same under every sharding
        decoder = SyntheticDecoder(pool_tr[:Bb], W, V, MAX_LENGTH, seed=7, device=dev, pool=ATT_POOL, per_row_noise=False)
        decoders = {}

        def load_batch(ids):
            idx = torch.tensor([i % P for i in ids], dtype=torch.long, device=dev)
            Tb = max(lengths[i] for i in ids)  # a length-sorted shard: the batch is as long as its longest utterance
            lg = pool_d.index_select(0, idx)[:, :Tb].contiguous()
            ln = torch.tensor([lengths[i] for i in ids], dtype=torch.long, device=dev)
            n = len(ids)
            dec = decoders.get(n)
            if dec is None:
                dec = decoder if n == Bb else SyntheticDecoder(pool_tr[:n], W, V, MAX_LENGTH, seed=7, device=dev, pool=ATT_POOL, per_row_noise=False)
                decoders[n] = dec
            dec.retarget([pool_tr[i % P] for i in ids])
            return lg, ln, dec

        def decode_batch(lg, ln, dec):
            proc = CTCRescorerLogitsProcessor(lg, ln, BLANK, EOS, 0, cfg.ctc_weight, W, -1, False, 1.0, materialize_state=False)
            return joint_beam_search_native(proc, dec, lg.shape[0], W, V, BOS, EOS, BLANK, max_length=MAX_LENGTH, device=dev, done_check_lag=1,
                                            fuse_topk=not args.no_fuse_topk)

        warm_ids = mine[: min(len(mine), Bb)]
        if warm_ids:  # warm-up: one batch, no collective
            sharding.decode_shard(warm_ids, Bb, load_batch, decode_batch, MAX_LENGTH, BLANK, dev)
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        ids, seqs, lens, scores = sharding.decode_shard(mine, Bb, load_batch, decode_batch, MAX_LENGTH, BLANK, dev)
        seqs, lens, scores = sharding.gather_hypotheses(ids, seqs, lens, scores, N, BLANK)
        e1.record()
        self.sync_all()
        wall = time.perf_counter() - t0
        ms = self.max_over_ranks(e0.elapsed_time(e1))[0]
        # every rank holds the whole result: check it against the aligned transcripts and fingerprint it
        want = torch.full((N, MAX_LENGTH), BLANK, dtype=torch.long)
        for i in range(N):
            t = pool_tr[i % P][:-1]
            want[i, : len(t)] = torch.tensor(t, dtype=torch.long)
        differing = int((seqs.cpu() != want).any(dim=1).sum())
        # copies of one utterance must decode identically whatever batch they were in
        copies_agree = bool((seqs.view(N // P, P, -1) == seqs[:P].unsqueeze(0)).all()) if N % P == 0 else None
        weights = torch.arange(1, MAX_LENGTH + 1, device=dev, dtype=torch.long).view(1, -1)
        checksum = int(((seqs * weights).sum(dim=1) % 1000003 * (torch.arange(1, N + 1, device=dev) % 1000003)).sum() % 1000000007)
        nb = (len(mine) + Bb - 1) // Bb
        return {"utterances": N, "value": N / (ms * 1e-3), "unit": UNIT, "ms": ms, "wall_s": wall, "scaling": "strong", "n_gpus": world,
                "utterances_this_rank": len(mine), "batches_this_rank": nb, "batch": Bb,
                "lengths": f"ragged (0.6 T .. T), {P} distinct utterances x {N // P}",
                "utterances_differing_from_aligned_transcript": differing, "copies_of_an_utterance_agree": copies_agree,
                "hypotheses_checksum": checksum,
                "note": ("sharding.shard_utterances -> decode_shard (native loop, batches trimmed to their longest utterance) -> "
                         "gather_hypotheses (the only collective: all_gather of the padded hypotheses); device-timed, max over ranks; "
                         "the checksum is over all hypotheses and must be identical at every N")}


class HFGenerate:
    """decode_fn for Runner.measure: transformers' generate() around a stub decoder (tests/hf_stub.py's idea on the GPU)."""

    def __init__(self, wl, runner):
        from transformers import GenerationMixin, PretrainedConfig, PreTrainedModel
        from transformers.modeling_outputs import CausalLMOutputWithPast

        from huggingface_asr_b200.generation import JointCTCAttentionGenerationMixin, joint_ctc_generation_config

        class _Cfg(PretrainedConfig):
            model_type = "ctcps_bench_stub_decoder"

            def __init__(self, vocab_size=64, **kw):
                super().__init__(**kw)
                self.vocab_size = vocab_size
                self.num_hidden_layers = 1

        class _Stub(JointCTCAttentionGenerationMixin, PreTrainedModel, GenerationMixin):
            config_class = _Cfg

            def __init__(self, config, decoder):
                super().__init__(config)
                self.dummy = torch.nn.Parameter(torch.zeros(1))
                self.decoder_fn, self.n = decoder, 0

            def forward(self, input_ids=None, attention_mask=None, past_key_values=None, use_cache=None, **kw):
                lp = self.decoder_fn(input_ids, self.n)
                self.n += 1
                return CausalLMOutputWithPast(logits=lp.unsqueeze(1), past_key_values=past_key_values)

            def generate(self, *a, **k):
                self.n = 0
                return super().generate(*a, **k)

        self.wl = wl
        self.model = _Stub(_Cfg(wl.V), wl.decoder).to(runner.dev)
        self.cfg = joint_ctc_generation_config(ctc_weight=wl.cfg.ctc_weight, num_beams=wl.W, max_length=MAX_LENGTH, pad_token_id=BLANK,
                                               eos_token_id=EOS, bos_token_id=BOS, do_sample=False, length_penalty=1.0, early_stopping=False,
                                               use_cache=True, num_return_sequences=1, return_dict_in_generate=True, output_scores=True)
        self.start = torch.full((wl.B, 1), BOS, dtype=torch.long, device=runner.dev)

    def __call__(self, lg, ln, timing):
        from huggingface_asr_b200.beam_search import BeamSearchOutput

        self.model.set_ctc_inputs(lg, ln)
        self.model.score_timing = timing  # the processor is built inside generate(): the mixin hands it the event list
        out = self.model.generate(self.start, generation_config=self.cfg)
        seqs = out.sequences[:, 1:]
        lengths = ((seqs != BLANK) & (seqs != EOS)).sum(dim=1)
        if seqs.shape[1] < MAX_LENGTH:
            seqs = torch.nn.functional.pad(seqs, (0, MAX_LENGTH - seqs.shape[1]), value=BLANK)
        scores = out.sequences_scores if getattr(out, "sequences_scores", None) is not None else torch.zeros(seqs.shape[0], device=seqs.device)
        return BeamSearchOutput(seqs, lengths, scores, int(self.model.n))


def run_ours(args):
    R = Runner(args)
    rank, world = R.rank, R.world
    wl = Workload(args.config, rank, R.dev, batch=args.batch)
    main_mode = args.state == "materialized"
    if args.state == "pre_beam":  # diagnostic / profiling runs of the N2 path only; the bench line is always a full-vocabulary mode
        if not args.profile:
            raise SystemExit("--state pre_beam is for --profile runs; the default run reports pre-beam under its own key")
        res = R.measure(wl, False, args.pre_beam)
    else:
        res = R.measure(wl, main_mode)
    if args.profile:
        if rank == 0:
            emit({"profile_run": True, "state": args.state, "ms_per_step": res["ms"] / res["steps"], "avg_score_ms": res["score_ms"]})
        return
    other = None if args.single_mode else R.measure(wl, not main_mode, clocks=False)
    # the same decode with every row scored at every step until the whole batch is finished and every frame from the prefix
    # length to T streamed (what the reference's dense scorer under HF's loop computes): the two exact shortcuts of the native
    # loop switched off -- same hypotheses bit for bit, and the scoring kernel's roofline on the full algorithmic bytes
    all_rows = None
    if not args.single_mode and main_mode is False and args.harness == "native":
        from huggingface_asr_b200 import _lib as _l

        prev_skip, prev_win = _l.lib().ctcps_set_skip_done(0), _l.lib().ctcps_set_frame_window(0)
        try:
            all_rows = R.measure(wl, False, steps=4, warm=2, e2e=False, clocks=False)
            all_rows_roof = R.roofline(wl, all_rows, False)
        finally:
            _l.lib().ctcps_set_skip_done(prev_skip)
            _l.lib().ctcps_set_frame_window(prev_win)
    pre = R.measure(wl, False, args.pre_beam, clocks=False) if (args.pre_beam > 0 and not args.single_mode) else None
    agreement = None
    key_full, key_pre = (wl.name, False, 0, args.harness), (wl.name, False, args.pre_beam, args.harness)
    if pre is not None and key_full in R.last_sequences:
        agreement = float((R.last_sequences[key_full] == R.last_sequences[key_pre]).all(dim=1).float().mean())

    hidden = None
    if args.hidden_dim > 0 and not args.single_mode:
        hidden = {"lazy": R.measure_from_hidden(wl, 0)}
        if pre is not None:
            hidden["pre_beam"] = R.measure_from_hidden(wl, args.pre_beam)
    head_roof = R.head_roofline(wl) if hidden is not None else None

    drop_in = None if (args.single_mode or args.no_drop_in) else R.drop_in(wl)

    m = R.summary(wl, res, main_mode)
    cfg_d = R.workload_config(wl, res, args.state)
    other_s = None if other is None else R.summary(wl, other, not main_mode)
    h2d_main, d2h_main, B_main = wl.h2d_bytes, wl.d2h_bytes, wl.B
    del wl  # frees 2 x 1.9 GB of device and pinned memory before the other configs are generated
    torch.cuda.empty_cache()

    extra_cfgs = {}
    if not args.single_mode and args.extra_configs:
        for name in [c for c in args.extra_configs.split(",") if c and c != args.config]:
            w2 = Workload(name, rank, R.dev)
            r_l = R.measure(w2, False, steps=6, warm=3, clocks=False)  # short decodes: a few more of them against host jitter
            entry = R.summary(w2, r_l, False)
            entry["config"] = R.workload_config(w2, r_l, "lazy")
            r_m = R.measure(w2, True, steps=2, warm=1, e2e=False, clocks=False)
            entry["materialized_state"] = R.summary(w2, r_m, True, with_e2e=False)
            if args.pre_beam > 0:
                r_p = R.measure(w2, False, args.pre_beam, steps=3, warm=2, e2e=False, clocks=False)
                entry["pre_beam"] = {"pre_beam_size": args.pre_beam, "value": world * w2.B * r_p["steps"] / (r_p["ms"] * 1e-3), "unit": UNIT,
                                     "ms_per_step": r_p["ms"] / r_p["steps"]}
            extra_cfgs[name] = entry
            del w2
            torch.cuda.empty_cache()

    c5 = None
    if not args.single_mode and args.c5_utterances > 0:
        c5 = R.c5_job()

    if rank == 0:
        cfg_d.update({
            "attention_scores": "SyntheticDecoder: log_softmax(noise + 10*onehot(transcript[n])) (the decoder is model code outside the path)",
            "max_length": MAX_LENGTH, "numa_bound_cpus": None if R.numa is None else len(R.numa),
            "l2": "inputs exceed L2: posteriors are 1.9 GB and (materialized) every scorer launch writes 8*T*BW*V bytes of state"})
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": res["steps"], "warmup": res["warm"],
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg_d, "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "roofline": m["roofline"],
            "clocks": m.get("clocks"),
        }
        if other_s is not None:
            line["lazy_state" if main_mode else "materialized_state"] = other_s
        if pre is not None:
            S = args.pre_beam
            n = world * B_main * pre["steps"]
            line["pre_beam"] = {
                "pre_beam_size": S, "value": n / (pre["ms"] * 1e-3), "unit": UNIT, "ms_per_step": pre["ms"] / pre["steps"],
                "e2e": {"value": n / (pre["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_main, "d2h_bytes_per_step": d2h_main},
                "gpu_launches": pre["launches"], "decode_steps_per_utterance_batch": pre["decode_steps"],
                "score_candidates_ms": pre["score_ms"], "one_best_agreement_with_full_vocabulary": agreement,
                "note": ("SURVEY 8(f) N2, not the reference's behaviour: only the top-S decoder tokens of every hypothesis are "
                         "CTC-scored (ESPnet pre-beam, S = 1.5 * beam by default), states selected with hyp*V+tok; sparse native "
                         "loop (no (BW,V) tensor); score_candidates_ms = CUDA-event time of ctcps_score_candidates"),
            }
        if all_rows is not None:
            line["every_row_every_frame"] = {
                "value": world * B_main * all_rows["steps"] / (all_rows["ms"] * 1e-3), "unit": UNIT, "ms_per_step": all_rows["ms"] / all_rows["steps"],
                "roofline": all_rows_roof,
                "note": "ctcps_set_skip_done(0) + ctcps_set_frame_window(0): the native loop scores the rows of finished utterances until the "
                        "whole batch is done and streams every frame from the prefix length to T, like the reference's dense scorer under HF's "
                        "beam search; same hypotheses, bit for bit (tests/test_gpu_fused_topk.py)"}
        if hidden is not None:
            d = args.hidden_dim
            cfg = CONFIGS[args.config]
            flops = 2.0 * B_main * cfg.T * cfg.V * d
            line["e2e_from_hidden"] = {
                k: {"value": world * B_main * args.steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_main,
                    "ctc_head_ms": head_ms, "ctc_head_tflops_fp32_equivalent": flops / (head_ms * 1e-3) / 1e12,
                    "ctc_head_implementation": impl, "transcripts_recovered": ok}
                for k, (ms, head_ms, h2d, ok, impl) in hidden.items()}
            if head_roof is not None:
                line["e2e_from_hidden"]["head_roofline"] = head_roof
            line["e2e_from_hidden"]["note"] = (
                f"SURVEY 8(f) N4 boundary, not the reference-facing call: host buffers hold the encoder hidden states (B,T,{d}) "
                "instead of the (B,T,V) logits; the CTC head (GEMM + log-softmax + padding) runs on the GPU inside the timed region; "
                "ctc_head_ms = CUDA-event time from the hidden states to the padded log-posteriors")
        if drop_in is not None:
            line["drop_in"] = drop_in
        if extra_cfgs:
            line["configs"] = extra_cfgs
        if c5 is not None:
            line["c5_job"] = c5
        if not args.no_cpu_baseline and world == 1:
            if R.affinity0 is not None:
                os.sched_setaffinity(0, R.affinity0)
            line["cpu_baseline"] = cpu_baseline(CONFIGS[args.config], args)
        emit(line)
    if R.dist is not None:
        R.dist.destroy_process_group()


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


def cpu_sample_size(cores):
    """Utterances of the bounded CPU sample: 8 on a host with >= 16 threads (about 25 s of oracle time per sample at C2)."""
    return 8 if cores >= 16 else (4 if cores >= 8 else 1)


def cpu_baseline(cfg, args):
    cores = os.cpu_count() or 1
    Bs = cpu_sample_size(cores)
    sec, steps = oracle_decode(cfg, Bs, seed=20240 + 2000)
    return {"value": Bs / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{Bs} utterance(s) of the {args.config} shape (T={cfg.T}, V={cfg.V}, beam {cfg.W}), full decode of {steps} steps, "
                      f"oracle/ctc_prefix_oracle.c with OpenMP on {cores} threads, {sec:.1f} s; utterances are independent, so utt/s of "
                      f"the sample is the rate of any batch size"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    warm = max(args.warmup, 0)
    # bounded sample: the whole --steps K --warmup W run should end within a few minutes.  The port decodes about
    # 0.02 utterances/s per host thread at C2 (0.33 utt/s on 16 threads); every step draws different utterances.
    est_rate = 0.02 * cores * (373.0 * 10 / (cfg.T * cfg.W))
    budget_s = 200.0
    Bs = args.batch or max(2 if cores >= 16 else 1, min(cpu_sample_size(cores), int(budget_s * est_rate / max(args.steps + warm, 1))))
    # same warm-up count as the GPU arm; a warm-up step is one decode of the same sample size
    for i in range(warm):
        oracle_decode(cfg, Bs, seed=1 + i)
    t = 0.0
    steps = 0
    for i in range(args.steps):
        sec, st = oracle_decode(cfg, Bs, seed=20240 + 2000 + i)
        t += sec
        steps += st
    val = Bs * args.steps / t
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)),
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg.name}", "utterances_per_gpu": cfg.B, "beam": cfg.W, "frames": cfg.T, "vocab": cfg.V,
                   "ctc_weight": cfg.ctc_weight, "logits": cfg.kind, "lengths": "equal (T)", "state": "materialized (the reference's data flow)",
                   "harness": "torch (beam_search.joint_beam_search, the loop shared with the GPU arm's drop_in.torch_harness)",
                   "decode_steps_per_utterance_batch": steps / args.steps, "max_length": MAX_LENGTH,
                   "sample_utterances_per_step": Bs,
                   "extrapolation": (f"a step decodes a bounded sample of {Bs} of the {cfg.B} utterances of the workload (same shape, same "
                                     "generator); utterances never interact, so the sample's utt/s is the CPU's rate on the full batch")},
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
    ap.add_argument("--single-mode", action="store_true", help="measure only --state on --config (no other keys)")
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
    ap.add_argument("--extra-configs", default="C1,C3,C4", help="other BASELINE configs measured in short runs under 'configs' ('' = none)")
    ap.add_argument("--c5-utterances", type=int, default=8192, help="utterances of the sharded C5 job under 'c5_job' (0 = skip)")
    ap.add_argument("--no-drop-in", action="store_true", help="skip the drop_in legs (processor under the torch harness and HF generate())")
    ap.add_argument("--no-fuse-topk", action="store_true",
                    help="native loop: write the dense joint scores and rank them in a second kernel (the round-1 step) instead of the fused per-tile top-2W")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--profile", action="store_true", help="for runs under ncu: honour a warm-up below 3 and skip the e2e/cpu legs (never a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
