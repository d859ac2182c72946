"""ctypes front-end of the CPU oracle (oracle/ctc_prefix_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package huggingface_asr_b200 never does.

Two layers:
  * plain functions over numpy arrays (log_softmax, pad, init_state, score, select, combine),
    one per C entry point, each the restatement of the reference lines cited in the C file;
  * OracleCTCPrefixScore / OracleCTCRescorerLogitsProcessor: the same scorer / processor
    surface as the reference (src/decoding/ctc_scorer.py:7,259) over torch CPU tensors, so the
    shared beam-search harness can drive the oracle exactly like the reference and the CUDA path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libctcps_oracle.so")
_lib = None

LOGZERO = -10000000000.0


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "ctc_prefix_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def set_threads(n: int) -> int:
    """Set the OpenMP team size of the oracle (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every core)."""
    L = lib()
    L.ctcps_oracle32_set_threads.restype = ctypes.c_int
    return int(L.ctcps_oracle32_set_threads(ctypes.c_int(int(n))))


def _dt(prec: int):
    return np.float32 if prec == 32 else np.float64


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _fn(name: str, prec: int):
    f = getattr(lib(), f"ctcps_oracle{prec}_{name}")
    f.restype = ctypes.c_int
    return f


def _i64(v):
    return ctypes.c_int64(int(v))


def log_softmax(logits: np.ndarray, prec: int = 32) -> np.ndarray:
    logits = np.ascontiguousarray(logits, dtype=_dt(prec))
    out = np.empty_like(logits)
    V = logits.shape[-1]
    _fn("log_softmax", prec)(_p(logits), _i64(logits.size // V), _i64(V), _p(out))
    return out


def pad(x: np.ndarray, lens, blank: int, prec: int = 32) -> np.ndarray:
    """In place on x (B,T,V), like the reference ctor."""
    assert x.dtype == _dt(prec) and x.flags.c_contiguous
    B, T, V = x.shape
    lens = np.ascontiguousarray(np.asarray(lens), dtype=np.int64)
    _fn("pad", prec)(_p(x), _p(lens), _i64(B), _i64(T), _i64(V), _i64(blank))
    return x


def init_state(x: np.ndarray, blank: int, W: int, prec: int = 32) -> np.ndarray:
    B, T, V = x.shape
    r0 = np.empty((T, 2, B * W), dtype=_dt(prec))
    _fn("init_state", prec)(_p(x), _i64(B), _i64(T), _i64(V), _i64(blank), _i64(W), _p(r0))
    return r0


def score(x, blank, r_prev, s_prev, last_ids, ol, W, scoring_ids=None, prec: int = 32, window=None):
    """Returns (token_scores, r, log_psi, idmap_or_None, early_return_flag).  window = (start, end) of ctc_scorer.py:127-136
    (None: max(ol, 1), T)."""
    B, T, V = x.shape
    BW = B * W
    dt = _dt(prec)
    r_prev = np.ascontiguousarray(r_prev, dtype=dt)
    assert r_prev.shape == (T, 2, BW), (r_prev.shape, (T, 2, BW))
    if s_prev is not None:
        s_prev = np.ascontiguousarray(np.broadcast_to(np.asarray(s_prev, dtype=dt), (BW, V)))
    last_ids = np.ascontiguousarray(np.asarray(last_ids), dtype=np.int64)
    S = 0
    idmap = None
    if scoring_ids is not None:
        scoring_ids = np.ascontiguousarray(np.asarray(scoring_ids), dtype=np.int64)
        S = scoring_ids.shape[-1]
        idmap = np.empty((BW, V), dtype=np.int64)
    snum = S if S > 0 else V
    r = np.empty((T, 2, BW, snum), dtype=dt)
    log_psi = np.empty((BW, V), dtype=dt)
    ts = np.empty((BW, V), dtype=dt)
    start, end = (max(ol, 1), T) if window is None else window
    early = _fn("score_window", prec)(
        _p(x), _i64(B), _i64(T), _i64(V), _i64(blank), _p(r_prev), _p(s_prev), _p(last_ids), _i64(ol), _i64(W),
        _p(scoring_ids), _i64(S), _i64(start), _i64(end), _p(r), _p(log_psi), _p(ts), _p(idmap))
    return ts, r, log_psi, idmap, bool(early)


def select(r, log_psi, best_ids, idmap, B, W, prec: int = 32):
    """Returns (r_new (T,2,BW), s_new (BW,))."""
    dt = _dt(prec)
    T = r.shape[0]
    V = log_psi.shape[1]
    S = 0 if idmap is None else r.shape[3]
    best_ids = np.ascontiguousarray(np.asarray(best_ids), dtype=np.int64).reshape(-1)
    r_new = np.empty((T, 2, B * W), dtype=dt)
    s_new = np.empty((B * W,), dtype=dt)
    _fn("select", prec)(_p(np.ascontiguousarray(r)), _p(np.ascontiguousarray(log_psi)), _p(best_ids), _p(idmap),
                        _i64(B), _i64(W), _i64(T), _i64(V), _i64(S), _p(r_new), _p(s_new))
    return r_new, s_new


def combine(scores, ctc, pad_id, w, apply_trick=False, eos=1, space=-1, trick_w=1.0, prec: int = 32):
    """In place on scores[:, pad]; returns next_token_scores."""
    dt = _dt(prec)
    assert scores.dtype == dt and scores.flags.c_contiguous
    ctc = np.ascontiguousarray(ctc, dtype=dt)
    BW, V = scores.shape
    out = np.empty_like(scores)
    real = ctypes.c_float if prec == 32 else ctypes.c_double
    _fn("combine", prec)(_p(scores), _p(ctc), _i64(BW), _i64(V), _i64(pad_id), real(w), ctypes.c_int(int(apply_trick)),
                         _i64(eos), _i64(space), real(trick_w), _p(out))
    return out


# --------------------------------------------------------------------------------------------
# Reference-shaped objects over torch CPU tensors (for the shared beam-search harness).
# --------------------------------------------------------------------------------------------
class OracleCTCPrefixScore:
    """Same surface as CTCPrefixScoreTH (ctc_scorer.py:7-207), the attention window of :127-136 included."""

    def __init__(self, x, xlens, blank, eos, margin=0, prec: int | None = None):
        import torch

        self._torch = torch
        self.prec = prec or (64 if x.dtype == torch.float64 else 32)
        self.logzero = LOGZERO
        self.blank, self.eos, self.margin = blank, eos, margin
        self.batch, self.input_length, self.odim = x.shape
        self.dtype = x.dtype
        self.device = torch.device("cpu")
        self._x = pad(np.ascontiguousarray(x.numpy()), np.asarray(xlens), blank, self.prec)
        self.end_frames = torch.as_tensor(xlens) - 1
        self.scoring_num = 0

    def __call__(self, y, state, scoring_ids=None, att_w=None):
        torch = self._torch
        ol = len(y[0]) - 1
        last_ids = np.asarray([int(yi[-1]) for yi in y], dtype=np.int64)
        n_bh = len(last_ids)
        W = n_bh // self.batch
        if state is None:
            r_prev = init_state(self._x, self.blank, W, self.prec)
            s_prev = None
            f_min_prev, f_max_prev = 0, 1                                   # :84-85
        else:
            r_prev, s_prev = state[0].numpy(), state[1].numpy()
            f_min_prev, f_max_prev = int(state[2]), int(state[3])
        sid = None if scoring_ids is None else scoring_ids.numpy()
        self.scoring_num = 0 if sid is None else sid.shape[-1]
        window, f_min, f_max = None, 0, 0
        if att_w is not None and self.margin > 0:                           # :127-132
            T = self.input_length
            f_arg = att_w.to(self.dtype) @ torch.arange(T, dtype=self.dtype)
            f_min = max(int(f_arg.min()), f_min_prev)
            f_max = max(int(f_arg.max()), f_max_prev)
            window = (min(f_max_prev, max(f_min - self.margin, ol, 1)), min(f_max + self.margin, T))
        ts, r, log_psi, idmap, _ = score(self._x, self.blank, r_prev, s_prev, last_ids, ol, W, sid, self.prec, window)
        return torch.from_numpy(ts), (torch.from_numpy(r), torch.from_numpy(log_psi), f_min, f_max,
                                      None if idmap is None else torch.from_numpy(idmap))

    def index_select_state(self, state, best_ids):
        torch = self._torch
        r, s, f_min, f_max, idmap = state
        n_bh = len(s)
        W = n_bh // self.batch
        r_new, s_new = select(r.numpy(), s.numpy(), best_ids.numpy(), None if idmap is None else idmap.numpy(),
                              self.batch, W, self.prec)
        s_new = torch.from_numpy(s_new).view(-1, 1).expand(n_bh, self.odim)
        return torch.from_numpy(r_new), s_new, f_min, f_max


    def extend_prob(self, x):
        """ctc_scorer.py:209-229: append frames; frames already seen keep their values; nothing is padded (xlens = [T'])."""
        torch = self._torch
        if self._x.shape[1] < x.shape[1]:
            new = np.ascontiguousarray(x.numpy()).astype(self._x.dtype, copy=True)
            new[:, : self._x.shape[1]] = self._x
            self._x = new
            self.input_length = x.shape[1]
            self.end_frames = torch.as_tensor([x.shape[1]]) - 1

    def extend_state(self, state):
        """ctc_scorer.py:231-256: extend the blank-only chain r_prev[t,1] = r_prev[t-1,1] + x[t,blank] to the new length."""
        torch = self._torch
        if state is None:
            return state
        r_prev, s_prev, f_min, f_max = state
        rp = r_prev.numpy()
        T = self.input_length
        out = np.full((T,) + rp.shape[1:], LOGZERO, dtype=rp.dtype)
        start = max(rp.shape[0], 1)
        out[: rp.shape[0]] = rp
        for t in range(start, T):
            out[t, 1] = out[t - 1, 1] + self._x[0, t, self.blank]  # single utterance, like the reference (:254)
        return torch.from_numpy(out), s_prev, f_min, f_max


class OracleCTCRescorerLogitsProcessor:
    """Same surface as CTCRescorerLogitsProcessor (ctc_scorer.py:259-354), on the oracle."""

    def __init__(self, encoder_logits, encoder_output_lens, pad_token_id, eos_token_id, ctc_margin, ctc_weight,
                 num_beams, space_token_id=-1, apply_eos_space_trick=False, eos_space_trick_weight=1.0, debug=False,
                 pre_beam_size=0, use_beam_idx=None):
        """pre_beam_size / use_beam_idx restate the policy ESPnet wraps around this scorer (espnet/nets/batch_beam_search.py:
        the top pre_beam_size tokens of the decoder scores of every hypothesis are handed to the scorer as scoring_ids;
        states are selected with source hypothesis * V + token); the reference's processor does neither (:326-330)."""
        import torch

        self._torch = torch
        prec = 64 if encoder_logits.dtype == torch.float64 else 32
        self.prec = prec
        self.pad_token_id = pad_token_id
        x = torch.from_numpy(log_softmax(encoder_logits.numpy(), prec))
        self.ctc_prefix_scorer = OracleCTCPrefixScore(x, encoder_output_lens, pad_token_id, eos_token_id, ctc_margin)
        self.ctc_weight, self.num_beams = ctc_weight, num_beams
        self.eos_token_id, self.space_token_id = eos_token_id, space_token_id
        self.apply_eos_space_trick, self.eos_space_trick_weight = apply_eos_space_trick, eos_space_trick_weight
        self.ctc_states = None
        self.pre_beam_size = int(pre_beam_size)
        self.use_beam_idx = (self.pre_beam_size > 0) if use_beam_idx is None else bool(use_beam_idx)
        self._best_ids = None

    def set_beam_idx(self, beam_idx):
        if self.use_beam_idx:
            W = self.num_beams
            self._best_ids = (beam_idx.view(-1, W) % W) * self.ctc_prefix_scorer.odim

    def __call__(self, input_ids, scores):
        torch = self._torch
        if self.ctc_states is not None:
            best = input_ids[:, -1].reshape(-1, self.num_beams)
            if self.use_beam_idx and self._best_ids is not None:
                best, self._best_ids = best + self._best_ids, None
            self.ctc_states = self.ctc_prefix_scorer.index_select_state(self.ctc_states, best)
        scoring_ids = None
        if self.pre_beam_size > 0:
            scores[:, self.pad_token_id] = LOGZERO  # :325, before the candidates are drawn
            scoring_ids = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, : self.pre_beam_size].contiguous()
        ctc_scores, self.ctc_states = self.ctc_prefix_scorer(input_ids, self.ctc_states, scoring_ids)
        s_np = scores.numpy()  # shares memory: the in-place scores[:, pad] = logzero reaches the caller
        assert s_np.flags.c_contiguous
        out = combine(s_np, ctc_scores.numpy(), self.pad_token_id, self.ctc_weight, self.apply_eos_space_trick,
                      self.eos_token_id, self.space_token_id, self.eos_space_trick_weight, self.prec)
        return torch.from_numpy(out)
