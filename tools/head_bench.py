"""CUDA-event timing of the CTC head (hidden states -> padded log-posteriors) at a BASELINE shape, both implementations
(diagnostic; bench.py's e2e_from_hidden.ctc_head_ms is the number on the record)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from huggingface_asr_b200.ctc_head import CTCHead  # noqa: E402
from huggingface_asr_b200.synthetic import BLANK, CONFIGS, make_encoder_hidden  # noqa: E402

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
d = 512
hid, w, b, lens, _ = make_encoder_hidden(cfg.B, cfg.T, cfg.V, d, seed=3)
hid, w, b, lens = hid.cuda(), w.cuda(), b.cuda(), lens.cuda()
flops = 2.0 * cfg.B * cfg.T * cfg.V * d
for impl in ("tcgen05", "cublas"):
    head = CTCHead(w, b, implementation=impl)
    for _ in range(3):
        head.log_posteriors(hid, lens, BLANK)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        head.log_posteriors(hid, lens, BLANK)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{impl:8s} hidden -> log-posteriors: {ms:.3f} ms  ({flops / ms / 1e9:.0f} TFLOP/s fp32-equivalent, {3 * flops / ms / 1e9:.0f} TFLOP/s of TF32 work)")
