"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/ctcps.h declares; argument
validation of the host layer; the product refuses CPU tensors; the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ctcps.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctcps_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from huggingface_asr_b200 import _lib

    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ctcps.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names


def test_no_compute_helpers_work_without_gpu():
    from huggingface_asr_b200 import _lib

    L = _lib.lib()
    assert L.ctcps_version() >= 100
    assert L.ctcps_padded_ld(5000) == 5056 and L.ctcps_padded_ld(5120) == 5120 and L.ctcps_padded_ld(1) == 64
    n = ctypes.c_size_t(0)
    assert L.ctcps_workspace_bytes(256, 373, 5000, 10, 0, ctypes.byref(n)) == 0
    assert 1_000_000 < n.value < 100_000_000
    assert L.ctcps_workspace_bytes(0, 373, 5000, 10, 0, ctypes.byref(n)) == -1
    assert b"bad sizes" in L.ctcps_error_string(-1)
    assert L.ctcps_error_string(0) == b"ok"


def test_argument_errors_are_reported_before_any_launch():
    from huggingface_asr_b200 import _lib

    L = _lib.lib()
    # null pointers / bad ids are rejected on the host, so this is safe without a GPU
    assert L.ctcps_init(None, 8, None, 1, 4, 8, 3, 1, None, 8, None, None) == -1
    assert L.ctcps_select(None, 8, None, None, None, 1, 1, 4, 8, 0, None, None, None) == -1
    with pytest.raises(ValueError, match="null pointer"):
        _lib.check(L.ctcps_score(None, 8, None, None, None, 0, 0, None, 0, 1, 1, 4, 8, 3, None, 0, None, None, 0.7, 0.3, None, 8,
                                 None, None, None, None, 0, None), "ctcps_score")


def test_argument_errors_of_the_pre_beam_and_step_entry_points():
    """Same for the N2 / N4 / native-step entry points: every one validates on the host before it launches anything."""
    import ctypes

    from huggingface_asr_b200 import _lib

    L = _lib.lib()
    assert L.ctcps_padded_lt(373) == 376 and L.ctcps_padded_lt(376) == 376
    assert L.ctcps_transpose_vt(None, 8, 1, 4, 8, None, 4, None) == -1
    assert L.ctcps_prebeam_topk(None, 4, 8, 3, 2, None, None, None) == -1
    assert L.ctcps_score_candidates(None, 4, None, None, None, 0, 1, 1, 4, 8, 3, None, 2, None, 0.7, 0.3, None, None, None, None, 0, 0,
                                    None) == -1
    assert L.ctcps_candidates_to_dense(None, None, None, None, None, None, 1, 8, 2, 0.7, 0.3, 0, 4, None, None, None, None) == -1
    assert L.ctcps_select_lazy_candidates(None, 4, None, None, None, 0, None, 2, None, None, 1, 1, 4, 8, None, None, None, 0, None) == -1
    assert L.ctcps_beam_step_candidates(None, None, 1, None, None, None, 8, 1, 1, 1, 8, 1, 3, 1.0, None, None, None, 8, None, None, 0,
                                        None, 0, 0, None, None) == -1
    assert L.ctcps_split_tf32(None, 4, 8, 0, None, None) == -1
    assert L.ctcps_decode_step(None, None, 0, None, None, None) == -1
    sess = _lib.DecodeSession()
    sess.S = 1  # neither full vocabulary nor >= 2 candidates
    dummy = ctypes.c_float(0.0)
    assert L.ctcps_decode_step(ctypes.byref(sess), ctypes.byref(dummy), 0, None, None, None) == -1
    assert L.ctcps_decode_finish(None, None) == -1
    # the session struct mirrors the C declaration (a mismatch would shift every pointer)
    assert ctypes.sizeof(_lib.DecodeSession) == L.ctcps_decode_session_size()


def test_cuda_library_is_sm_100a_with_tma():
    from huggingface_asr_b200 import _lib

    _lib.lib()
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTMALDG" in sass, "the recursion kernel must stage x tiles with TMA (cp.async.bulk.tensor)"
    assert "UBLKCP" in sass
    # tensor cores: only the CTC head contracts -- tcgen05.mma with TMEM accumulators (UTCHMMA / LDTM), never
    # the legacy mma.sync (a bare HMMA); the prefix-scoring kernels are HBM-bound FP32 streams and must stay off them
    per_fn = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_fn[name] = []
        elif name is not None:
            per_fn[name].append(line)
    tc = {fn for fn, body in per_fn.items() if any("UTCHMMA" in ln for ln in body)}
    assert tc and all("k_head_gemm" in fn for fn in tc), tc
    assert any("LDTM" in ln for fn in tc for ln in per_fn[fn]), "the head's epilogue reads its accumulators from TMEM"
    assert not any(re.search(r"(?<!UTC)HMMA", ln) for body in per_fn.values() for ln in body), "no mma.sync kernels"


def test_scorer_refuses_cpu_tensors_and_bad_dtypes():
    from huggingface_asr_b200.decoding.ctc_scorer import CTCPrefixScoreTH, CTCRescorerLogitsProcessor, LogSoftmaxProcessor

    with pytest.raises(RuntimeError, match="no CPU path"):
        CTCPrefixScoreTH(torch.zeros(1, 4, 8), torch.tensor([4]), 3, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        CTCRescorerLogitsProcessor(torch.zeros(1, 4, 8), torch.tensor([4]), 3, 1, 0, 0.3, 2, -1, False, 1.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        LogSoftmaxProcessor()(None, torch.zeros(2, 8))
    with pytest.raises(TypeError):
        CTCPrefixScoreTH([[0.0]], [1], 3, 1)


def test_processor_is_a_transformers_logits_processor():
    from transformers import LogitsProcessor

    from huggingface_asr_b200.decoding import ctc_scorer

    assert issubclass(ctc_scorer.CTCRescorerLogitsProcessor, LogitsProcessor)
    assert issubclass(ctc_scorer.LogSoftmaxProcessor, LogitsProcessor)
    import inspect

    sig = inspect.signature(ctc_scorer.CTCRescorerLogitsProcessor.__init__)
    # the reference's positional signature, then our keyword-only extension
    assert list(sig.parameters)[1:12] == ["encoder_logits", "encoder_output_lens", "pad_token_id", "eos_token_id", "ctc_margin",
                                          "ctc_weight", "num_beams", "space_token_id", "apply_eos_space_trick",
                                          "eos_space_trick_weight", "debug"]
    extra = list(sig.parameters.values())[12:]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in extra)
    sig = inspect.signature(ctc_scorer.CTCPrefixScoreTH.__init__)
    assert list(sig.parameters)[1:] == ["x", "xlens", "blank", "eos", "margin"]
    for m in ("__call__", "index_select_state", "extend_prob", "extend_state"):
        assert callable(getattr(ctc_scorer.CTCPrefixScoreTH, m))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "huggingface_asr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # no import, no dlopen / ctypes load, no path into oracle/: comments may mention the oracle, code may not reach it
                for needle in ("from oracle", "import oracle", "oracle/", "oracle.", "libctcps_oracle", "oracle/_build"):
                    assert needle not in src, f"{os.path.relpath(os.path.join(dirpath, f), ROOT)} reaches into the oracle ({needle!r})"
    code = "import sys; import huggingface_asr_b200.decoding.ctc_scorer, huggingface_asr_b200.beam_search, huggingface_asr_b200.sharding; " \
           "assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported by the product'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_reference_frame_counts():
    from huggingface_asr_b200.synthetic import frames_for_seconds

    assert [frames_for_seconds(s) for s in (10, 15, 30)] == [248, 373, 748]
    assert [frames_for_seconds(s, 1) for s in (10, 15, 30)] == [250, 375, 750]


def test_ab_switch_query_set_restore():
    """ctcps_set_select_pscan: any argument other than 0 / 1 only queries; the default is the time-parallel kernel."""
    from huggingface_asr_b200 import _lib

    L = _lib.lib()
    fn = L.ctcps_set_select_pscan
    prev = fn(-1)
    assert prev in (0, 1)
    assert fn(1) == prev and fn(7) == 1 and fn(0) == 1 and fn(-1) == 0
    fn(prev)
    import os

    if "CTCPS_SELECT_PSCAN" not in os.environ:
        assert L.ctcps_set_select_pscan(-1) == 1
