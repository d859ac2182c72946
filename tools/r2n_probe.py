"""r2n probe: why is the full-vocabulary decode slower after the tcgen05 head?  (diagnostic)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from huggingface_asr_b200.beam_search import joint_beam_search_native, resolve_score_timing  # noqa: E402
from huggingface_asr_b200.ctc_head import CTCHead  # noqa: E402
from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor  # noqa: E402
from huggingface_asr_b200.synthetic import BLANK, BOS, CONFIGS, EOS, SyntheticDecoder, make_encoder_hidden  # noqa: E402

dev = torch.device("cuda")
cfg = CONFIGS["C2"]
B, W, T, V, d = cfg.B, cfg.W, cfg.T, cfg.V, 512
hid, w, b, lens, tr = make_encoder_hidden(B, T, V, d, seed=3)
hid, w, b, lens = hid.to(dev), w.to(dev), b.to(dev), lens.to(dev)
dec = SyntheticDecoder(tr, W, V, 128, seed=7, device=dev)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for impl in ("cublas", "tcgen05", "cublas", "tcgen05"):
    head = CTCHead(w, b, implementation=impl)
    for rep in range(3):
        torch.cuda.synchronize()
        e0 = ev()
        proc = CTCRescorerLogitsProcessor.from_encoder_hidden_states(hid, head, lens, BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
        e1 = ev()
        st = []
        out = joint_beam_search_native(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=128, device=dev, done_check_lag=1, score_timing=st)
        e2 = ev()
        torch.cuda.synchronize()
        ms = resolve_score_timing(st)
        x = proc.ctc_prefix_scorer._x
        print(f"{impl:8s} rep {rep}: head {e0.elapsed_time(e1):6.2f} ms, decode {e1.elapsed_time(e2):6.2f} ms ({out.steps} steps), scoring call avg {sum(ms) / len(ms):.4f} ms "
              f"min {min(ms):.4f} max {max(ms):.4f}; x ptr % 1024 = {x.data_ptr() % 1024}, stride {x.stride()}, finite {bool(torch.isfinite(x[..., :V]).all())}", flush=True)
    del head
