"""r2f probes (diagnostic, not bench values): (1) pre-beam decode time per decode at C3 / C2 / C1 with the native loop;
(2) the torch-harness drop-in leg resident vs with a concurrent H2D copy."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from huggingface_asr_b200.beam_search import joint_beam_search, joint_beam_search_native  # noqa: E402
from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor  # noqa: E402
from huggingface_asr_b200.synthetic import BLANK, BOS, CONFIGS, EOS, SyntheticDecoder, make_encoder_logits  # noqa: E402

dev = torch.device("cuda")


def timed(fn, n=4):
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    return out, r


for name in ("C3", "C2", "C1"):
    cfg = CONFIGS[name]
    B, W, T, V = cfg.B, cfg.W, cfg.T, cfg.V
    lg, ln, tr = make_encoder_logits(B, T, V, cfg.kind, False, seed=5)
    lg, ln = lg.to(dev), ln.to(dev)
    dec = SyntheticDecoder(tr, W, V, 128, seed=7, device=dev)
    for S in (15, 0):
        def run():
            proc = CTCRescorerLogitsProcessor(lg, ln, BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False, pre_beam_size=S)
            return joint_beam_search_native(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=128, device=dev, done_check_lag=1)
        ms, o = timed(run, 5)
        print(f"{name} S={S:2d} native: " + " ".join(f"{m:7.2f}" for m in ms) + f" ms, steps {o.steps}", flush=True)
    if name == "C2":
        def run_t():
            proc = CTCRescorerLogitsProcessor(lg, ln, BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
            return joint_beam_search(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=128, device=dev)
        ms, o = timed(run_t, 4)
        print("C2 torch harness alone:        " + " ".join(f"{m:7.2f}" for m in ms), flush=True)
        host = lg.cpu().pin_memory()
        buf = torch.empty_like(lg)
        cs = torch.cuda.Stream()

        def run_c():
            with torch.cuda.stream(cs):
                buf.copy_(host, non_blocking=True)
            return run_t()
        ms, o = timed(run_c, 4)
        print("C2 torch harness + H2D copy:   " + " ".join(f"{m:7.2f}" for m in ms), flush=True)
        t0 = time.perf_counter()
        for _ in range(50):
            a = torch.empty((2560, 5000), device=dev)
        torch.cuda.synchronize()
        print("50 x torch.empty(51 MB):", (time.perf_counter() - t0) * 1e3, "ms")
    del lg, dec
    torch.cuda.empty_cache()
