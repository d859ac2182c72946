"""Where a decode step goes when the processor is used as a drop-in (diagnostic, prints one JSON line):
CUDA-event time and host time of CTCRescorerLogitsProcessor.__call__ against the whole step, under the torch harness
(beam_search.joint_beam_search: HF's beam search restated in torch) at a BASELINE shape.
    python tools/dropin_probe.py [C2]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from huggingface_asr_b200.beam_search import joint_beam_search  # noqa: E402
from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor  # noqa: E402
from huggingface_asr_b200.synthetic import BLANK, BOS, CONFIGS, EOS, SyntheticDecoder, make_encoder_logits  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = CONFIGS[name]
logits, lens, transcripts = make_encoder_logits(cfg.B, cfg.T, cfg.V, cfg.kind, False, seed=4)
logits, lens = logits.cuda(), lens.cuda()
dec = SyntheticDecoder(transcripts, cfg.W, cfg.V, 128, seed=7, device="cuda")


class Timed:
    def __init__(self, proc):
        self.proc, self.events, self.host = proc, [], 0.0

    def __getattr__(self, k):
        return getattr(self.proc, k)

    def __call__(self, ids, scores):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = self.proc(ids, scores)
        e1.record()
        self.host += time.perf_counter() - t0
        self.events.append((e0, e1))
        return out


res = {}
for it in range(3):
    proc = Timed(CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, 0.3, cfg.W, -1, False, 1.0))
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    out = joint_beam_search(proc, dec, cfg.B, cfg.W, cfg.V, BOS, EOS, BLANK, max_length=128, device="cuda")
    f1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    call_ms = sum(a.elapsed_time(b) for a, b in proc.events)
    res = {"config": name, "steps": out.steps, "decode_ms": f0.elapsed_time(f1), "wall_ms": wall * 1e3, "processor_call_gpu_ms": call_ms,
           "processor_call_host_ms": proc.host * 1e3, "processor_share_of_decode": call_ms / f0.elapsed_time(f1),
           "per_step_us": {"decode": f0.elapsed_time(f1) / out.steps * 1e3, "processor_call_gpu": call_ms / out.steps * 1e3,
                           "processor_call_host": proc.host / out.steps * 1e6}}
print(json.dumps(res))
