"""The CTC head in front of the scorer: logits = hidden @ W^T + b at fp32 accuracy on tensor cores.

In the reference the head is the stock `Wav2Vec2ForCTC.lm_head` Linear (src/reguler/e_branchformer.py:245-252), an fp32
SGEMM of (B*T, d) x (d, V) on the CUDA cores (C2: 0.49 TFLOP, ~7 ms on a B200) whose (B,T,V) output the processor then
log-softmaxes.  A single-pass TF32 GEMM would be 10x faster but misses the 1e-4 log-space tolerance of the path
(10-bit mantissas).  Here both operands are split into TF32-exact high and low parts by `ctcps_split_tf32` and stacked
along K, so ONE TF32 GEMM with fp32 accumulation computes hi*lo + lo*hi + hi*hi -- the dropped lo*lo term is 2^-22
relative -- at a third of the TF32 rate.  The GEMM itself is a plain library call (cuBLAS through
torch.addmm, bias in its epilogue); the log-softmax + padding that follows is K-a (`ctcps_init`).

Status: N4 is only started -- fusing the bias / row-max / sum-exp epilogue into a hand-written tcgen05 GEMM (so that the
logits are never written to HBM) is the remaining work.
"""
from __future__ import annotations

import contextlib

import torch

from . import _lib


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


@contextlib.contextmanager
def _tf32_matmul():
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def split_tf32(x: torch.Tensor, weight_order: bool) -> torch.Tensor:
    """(n,d) fp32 -> (n,3d): [hi | lo | hi] for activations, [lo | hi | hi] for weights (see ctcps_split_tf32)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or not x.is_contiguous():
        raise ValueError("split_tf32 needs a contiguous 2-D float32 CUDA tensor")
    n, d = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((n, 3 * d), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().ctcps_split_tf32(x.data_ptr(), n, d, int(weight_order), out.data_ptr(), _stream(x.device)),
                   "ctcps_split_tf32")
    return out


class CTCHead:
    """`lm_head` of the encoder: weight (V,d), bias (V) or None.  The weight is split once; __call__ maps encoder hidden
    states (B,T,d) to the (B,T,V) logits the processor takes."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor | None = None):
        if weight.dim() != 2 or weight.shape[1] % 4 != 0:
            raise ValueError("lm_head weight must be (V, d) with d a multiple of 4")
        self.vocab, self.dim = (int(v) for v in weight.shape)
        self.weight3 = split_tf32(weight.detach().contiguous(), weight_order=True)
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous()

    @torch.no_grad()
    def __call__(self, hidden: torch.Tensor) -> torch.Tensor:
        if hidden.dim() != 3 or hidden.shape[-1] != self.dim:
            raise ValueError(f"hidden states must be (B, T, {self.dim}), got {tuple(hidden.shape)}")
        B, T, d = hidden.shape
        h3 = split_tf32(hidden.reshape(B * T, d).contiguous(), weight_order=False)
        with _tf32_matmul():
            if self.bias is not None:
                logits = torch.addmm(self.bias, h3, self.weight3.t())
            else:
                logits = torch.mm(h3, self.weight3.t())
        return logits.view(B, T, self.vocab)
