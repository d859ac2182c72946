"""Builds and loads libctcps_b200.so (the C ABI of include/ctcps.h) through ctypes.

There is deliberately no fallback: if the library is missing and cannot be built, or a tensor is
not on a CUDA device, the callers raise.  Python and torch are plumbing (device memory, streams);
every kernel on the path lives in csrc/ (ctcps_kernels.cu + the .cuh files it includes).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SRC = os.path.join(_PKG, "csrc", "ctcps_kernels.cu")
SOURCES = [SRC, os.path.join(_PKG, "csrc", "ctcps_head.cu")]  # translation units of the one library
INCLUDE = os.path.join(_ROOT, "include")
LIB_PATH = os.path.join(_PKG, "libctcps_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

_lib = None


class CtcpsError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise CtcpsError("nvcc not found: cannot build libctcps_b200.so (there is no CPU fallback)")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ctcps_kernels.cu for sm_100a into the package directory (in-tree)."""
    import fcntl

    csrc = os.path.dirname(SRC)
    deps = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh"))] + [os.path.join(INCLUDE, "ctcps.h")]

    def fresh():
        return os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps)

    if not force and fresh():
        return LIB_PATH
    # several ranks of one torchrun job may get here together: one builds, the others wait and reuse
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return LIB_PATH
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, *SOURCES, "-o", tmp]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            env = dict(os.environ)
            env.pop("CC", None)  # the image exports a gcc wrapper nvcc must not be pointed at
            env.pop("CXX", None)
            res = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise CtcpsError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
            os.replace(tmp, LIB_PATH)  # atomic: a concurrent loader sees the old or the new file, never a partial one
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_i, _i64, _f, _p, _sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# symbol -> argtypes; must list every function include/ctcps.h declares (tests check this)
SIGNATURES = {
    "ctcps_version": [],
    "ctcps_error_string": [_i],
    "ctcps_padded_ld": [_i],
    "ctcps_set_select_pscan": [_i],
    "ctcps_set_psi_prefetch": [_i],
    "ctcps_set_skip_done": [_i],
    "ctcps_set_frame_window": [_i],
    "ctcps_set_stream_counter": [_p],
    "ctcps_stream_chunk_bytes": [_i],
    "ctcps_set_psi_max_group": [_i],
    "ctcps_workspace_bytes": [_i, _i, _i, _i, _i, ctypes.POINTER(_sz)],
    "ctcps_init": [_p, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p, _p],
    "ctcps_log_softmax": [_p, _i, _p, _i, _i, _i, _p],
    "ctcps_initial_state": [_p, _i, _i, _i, _i, _p, _p],
    "ctcps_score": [_p, _i, _p, _p, _p, _i64, _i64, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _f, _f, _p, _i, _p, _p, _p,
                    _p, _sz, _p],
    "ctcps_score_window": [_p, _i, _p, _p, _p, _i64, _i64, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _f, _f, _p, _i, _p, _p,
                           _p, _p, _sz, _p],
    "ctcps_score_lazy": [_p, _i, _p, _p, _p, _i64, _i64, _p, _i, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _sz, _i, _p],
    "ctcps_score_lazy_lens": [_p, _i, _p, _p, _p, _p, _i64, _i64, _p, _i, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _sz, _i, _p],
    "ctcps_select_lazy": [_p, _i, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p],
    "ctcps_topk_lists_shape": [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)],
    "ctcps_score_lazy_topk": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _sz, _i, _p],
    "ctcps_score_lazy_topk_active": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _p, _p, _p, _p, _sz, _i, _p],
    "ctcps_beam_step_lists": [_p, _i, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _i64, _p, _p, _sz, _p, _i, _i64, _p, _p,
                              _p],
    "ctcps_beam_step_workspace_bytes": [_i, _i, ctypes.POINTER(_sz)],
    "ctcps_beam_step": [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _i64, _p, _p, _sz, _p, _i, _i64, _p, _p],
    "ctcps_padded_lt": [_i],
    "ctcps_transpose_vt": [_p, _i, _i, _i, _i, _p, _i, _p],
    "ctcps_prebeam_topk": [_p, _i, _i, _i, _i, _p, _p, _p],
    "ctcps_score_candidates": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _f, _f, _p, _p, _p, _p, _sz, _i, _p],
    "ctcps_candidates_to_dense": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _i, _i, _p, _p, _p, _p],
    "ctcps_select_lazy_candidates": [_p, _i, _p, _p, _p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p],
    "ctcps_async_create": [ctypes.POINTER(_p), ctypes.POINTER(_p), ctypes.POINTER(_p)],
    "ctcps_async_destroy": [_p, _p, _p],
    "ctcps_event_create": [ctypes.POINTER(_p)],
    "ctcps_event_destroy": [_p],
    "ctcps_event_elapsed_ms": [_p, _p, ctypes.POINTER(_f)],
    "ctcps_decode_step": [_p, _p, _i, _p, _p, _p],
    "ctcps_decode_finish": [_p, _p],
    "ctcps_decode_session_size": [],
    "ctcps_split_tf32": [_p, _i64, _i, _i, _p, _p],
    "ctcps_head_workspace_bytes": [_i64, _i, ctypes.POINTER(_sz)],
    "ctcps_split_hi_lo": [_p, _i64, _p, _p, _p],
    "ctcps_head_weight_bytes": [_i, _i, ctypes.POINTER(_sz)],
    "ctcps_head_prepare_weight": [_p, _i, _i, _p, _p, _p],
    "ctcps_ctc_head": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _sz, _p],
    "ctcps_beam_step_candidates": [_p, _p, _i, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _i64, _p, _p, _sz, _p, _i,
                                   _i64, _p, _p],
    "ctcps_select": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p],
    "ctcps_eos_space_trick": [_p, _p, _p, _i, _i, _i, _i, _f, _p],
}


class DecodeSession(ctypes.Structure):
    """ctcps_decode_session of include/ctcps.h (same field order and types)."""
    _fields_ = [
        ("B", ctypes.c_int32), ("W", ctypes.c_int32), ("T", ctypes.c_int32), ("V", ctypes.c_int32), ("S", ctypes.c_int32),
        ("blank", ctypes.c_int32), ("eos", ctypes.c_int32), ("pad", ctypes.c_int32), ("use_beam_idx", ctypes.c_int32),
        ("ldx", ctypes.c_int32), ("ldt", ctypes.c_int32), ("ring", ctypes.c_int32),
        ("one_minus_w", _f), ("w", _f), ("length_penalty", _f),
        ("x_logp", _p), ("x_vt", _p), ("blank_lp", _p), ("r0", _p),
        ("r_sel", _p * 2), ("s_sel", _p * 2), ("last_ids", _p * 2), ("cand_ids", _p * 2), ("cand_att", _p * 2),
        ("cand_log_psi", _p * 2), ("cand_joint", _p), ("log_psi", _p * 2), ("joint", _p),
        ("score_ws", _p), ("score_ws_bytes", _sz),
        ("beam_scores", _p), ("ids", _p * 2), ("ld_ids", _i64), ("pool_scores", _p), ("pool_lens", _p), ("pool_seqs", _p),
        ("ld_pool", _i64), ("done", _p), ("beam_ws", _p), ("beam_ws_bytes", _sz), ("done_ring", _p), ("best_ids", _p),
        ("side_stream", _p), ("ev_step", _p), ("ev_select", _p),
        ("tile_lists", _p), ("tag_base", _i64), ("xlens", _p),
    ]


def lib() -> ctypes.CDLL:
    """The loaded library; builds it on first use if the .so is absent or stale."""
    global _lib
    if _lib is None:
        path = build()
        L = ctypes.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = (ctypes.c_char_p if name == "ctcps_error_string" else
                          ctypes.c_size_t if name == "ctcps_decode_session_size" else ctypes.c_int)
        if L.ctcps_decode_session_size() != ctypes.sizeof(DecodeSession):
            raise CtcpsError("DecodeSession (huggingface_asr_b200/_lib.py) does not mirror ctcps_decode_session (include/ctcps.h)")
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ctcps_error_string(rc).decode()
        kind = ValueError if rc < 0 else CtcpsError
        raise kind(f"{what}: {msg} (code {rc})")


def launches_per_score(S: int = 0) -> int:
    """Kernel launches one ctcps_score call makes (for bench.py's gpu_launches count)."""
    return 2 if S == 0 else 5
