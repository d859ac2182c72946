"""The CTC head in front of the scorer: hidden states -> log-posteriors on the tensor cores at fp32 accuracy.

In the reference the head is the stock `Wav2Vec2ForCTC.lm_head` Linear (src/reguler/e_branchformer.py:245-252), an fp32
SGEMM of (B*T, d) x (d, V) on the CUDA cores (C2: 0.49 TFLOP, ~7 ms on a B200) whose (B,T,V) output the processor then
log-softmaxes (src/decoding/ctc_scorer.py:279).  A single-pass TF32 GEMM would be 10x faster but misses the 1e-4 log-space
tolerance of the path (10-bit mantissas).  Both operands are therefore split into two 11-bit parts and three tensor-core
products -- hi*lo + lo*hi + hi*hi, the dropped lo*lo term is 2^-22 relative -- give fp32-grade logits.

Two implementations behind one class (`implementation`, env CTCPS_HEAD):
  "tcgen05" (default)  csrc/ctcps_head.cu: a hand-written sm_100a kernel.  The parts are fp16 (an fp16 significand has TF32's 11
                       bits in half the bytes at twice the MMA rate; rows of the hidden states and the weight are scaled by powers
                       of two into fp16's range first), fed by TMA bulk copies to `tcgen05.mma.kind::f16`, the large product and
                       the two small cross terms in separate TMEM accumulators, bias and the row-wise softmax statistics in the
                       TMEM -> register drain, logits written by TMA stores straight into the scorer's padded posterior buffer --
                       followed by one streaming normalisation pass (log-softmax, length padding, blank column).
  "cublas"             round 1's form, kept for A/B: `ctcps_split_tf32` stacks the split operands along K and ONE library TF32
                       GEMM (cuBLAS through torch.addmm) accumulates the three products; K-a (`ctcps_init`) follows.
"""
from __future__ import annotations

import contextlib
import ctypes
import os

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


def split_hi_lo(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """fp32 tensor -> (TF32-exact high part, fp32 remainder), same shape (ctcps_split_hi_lo)."""
    if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous() or x.numel() % 4:
        raise ValueError("split_hi_lo needs a contiguous float32 CUDA tensor whose size is a multiple of 4")
    with torch.cuda.device(x.device):
        hi, lo = torch.empty_like(x), torch.empty_like(x)
        _lib.check(_lib.lib().ctcps_split_hi_lo(x.data_ptr(), x.numel(), hi.data_ptr(), lo.data_ptr(), _stream(x.device)), "ctcps_split_hi_lo")
    return hi, lo


class CTCHead:
    """`lm_head` of the encoder: weight (V,d), bias (V) or None.  The weight is split once; `log_posteriors` maps encoder
    hidden states (B,T,d) to what the scorer keeps -- padded log-posteriors (B,T,ldx) and the blank column -- and `__call__` to
    the (B,T,V) logits (the head alone)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor | None = None, implementation: str | None = None):
        if weight.dim() != 2 or weight.shape[1] % 4 != 0:
            raise ValueError("lm_head weight must be (V, d) with d a multiple of 4")
        if not weight.is_cuda:
            raise RuntimeError("the CTC head runs on the GPU only (there is no CPU path)")
        self.vocab, self.dim = (int(v) for v in weight.shape)
        impl = implementation or os.environ.get("CTCPS_HEAD", "tcgen05")
        if impl not in ("tcgen05", "cublas"):
            raise ValueError(f"unknown CTC head implementation {impl!r}")
        if impl == "tcgen05" and self.dim % 16 != 0:
            raise ValueError("the tcgen05 CTC head needs a hidden size that is a multiple of 16")
        self.implementation = impl
        w = weight.detach().to(torch.float32).contiguous()
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous()
        if impl == "tcgen05":
            # the weight, split and laid out tile by tile in the swizzled image the kernel's bulk copies expect: once per model
            nb = ctypes.c_size_t(0)
            _lib.check(_lib.lib().ctcps_head_weight_bytes(self.vocab, self.dim, ctypes.byref(nb)), "ctcps_head_weight_bytes")
            with torch.cuda.device(w.device):
                self.w_hi = torch.empty((nb.value // 4,), dtype=torch.float32, device=w.device)
                self.w_lo = torch.empty((nb.value // 4,), dtype=torch.float32, device=w.device)
                _lib.check(_lib.lib().ctcps_head_prepare_weight(w.data_ptr(), self.vocab, self.dim, self.w_hi.data_ptr(), self.w_lo.data_ptr(),
                                                                _stream(w.device)), "ctcps_head_prepare_weight")
            self._ws = None
        else:
            self.weight3 = split_tf32(w, weight_order=True)

    def _workspace(self, n: int, device):
        need = ctypes.c_size_t(0)
        _lib.check(_lib.lib().ctcps_head_workspace_bytes(n, self.dim, ctypes.byref(need)), "ctcps_head_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != device:
            self._ws = torch.empty((need.value,), dtype=torch.uint8, device=device)
        return self._ws

    def _check_hidden(self, hidden):
        if hidden.dim() != 3 or hidden.shape[-1] != self.dim:
            raise ValueError(f"hidden states must be (B, T, {self.dim}), got {tuple(hidden.shape)}")
        if not hidden.is_cuda:
            raise RuntimeError("hidden states must be a CUDA tensor (there is no CPU path)")
        if hidden.is_floating_point() and hidden.dtype != torch.float32:
            hidden = hidden.float()
        return hidden.contiguous()

    def _run_tcgen05(self, hidden, lens, blank, out, ldx, blank_lp, apply_log_softmax):
        B, T, _ = hidden.shape
        L = _lib.lib()
        with torch.cuda.device(hidden.device):
            ws = self._workspace(B * T, hidden.device)
            _lib.check(L.ctcps_ctc_head(hidden.data_ptr(), self.w_hi.data_ptr(), self.w_lo.data_ptr(),
                                        None if self.bias is None else self.bias.data_ptr(), None if lens is None else lens.data_ptr(),
                                        B, T, self.dim, self.vocab, int(blank), int(apply_log_softmax), out.data_ptr(), ldx,
                                        None if blank_lp is None else blank_lp.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _stream(hidden.device)), "ctcps_ctc_head")

    @torch.no_grad()
    def log_posteriors(self, hidden: torch.Tensor, lens: torch.Tensor, blank: int):
        """(x (B,T,ldx) padded log-posteriors, blank_lp (B,T)): what CTCPrefixScoreTH.from_logits builds from the logits."""
        hidden = self._check_hidden(hidden)
        B, T, _ = hidden.shape
        L = _lib.lib()
        ldx = L.ctcps_padded_ld(self.vocab)
        lens = lens.to(device=hidden.device, dtype=torch.long).contiguous()
        with torch.cuda.device(hidden.device):
            x = torch.empty((B, T, ldx), dtype=torch.float32, device=hidden.device)
            blank_lp = torch.empty((B, T), dtype=torch.float32, device=hidden.device)
            if self.implementation == "tcgen05":
                self._run_tcgen05(hidden, lens, blank, x, ldx, blank_lp, True)
            else:
                logits = self(hidden)
                _lib.check(L.ctcps_init(logits.data_ptr(), self.vocab, lens.data_ptr(), B, T, self.vocab, int(blank), 1, x.data_ptr(), ldx,
                                        blank_lp.data_ptr(), _stream(hidden.device)), "ctcps_init")
        return x, blank_lp

    @torch.no_grad()
    def __call__(self, hidden: torch.Tensor) -> torch.Tensor:
        """The head alone: (B,T,V) logits."""
        hidden = self._check_hidden(hidden)
        B, T, d = hidden.shape
        if self.implementation == "tcgen05":
            ldz = (self.vocab + 3) & ~3
            with torch.cuda.device(hidden.device):
                z = torch.empty((B, T, ldz), dtype=torch.float32, device=hidden.device)
            self._run_tcgen05(hidden, None, 0, z, ldz, None, False)
            return z[..., : self.vocab]
        h3 = split_tf32(hidden.reshape(B * T, d), weight_order=False)
        with _tf32_matmul():
            if self.bias is not None:
                logits = torch.addmm(self.bias, h3, self.weight3.t())
            else:
                logits = torch.mm(h3, self.weight3.t())
        return logits.view(B, T, self.vocab)
