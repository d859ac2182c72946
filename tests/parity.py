"""Shared parity machinery: replay the committed golden vectors through ANY implementation that
exposes the reference's scorer / processor surface (oracle on CPU, sm_100a scorer on GPU).

Parity criterion (BASELINE.json north_star):
  * entries where the reference is > -1e9: |new - ref| <= ATOL (1e-4 absolute, fp32 log space)
    plus RTOL * |ref| -- the relative term only matters for |ref| >~ 100, where one fp32 ulp is
    already 8e-6 .. 6e-5 and two correct fp32 implementations differ by a few ulps;
  * entries where the reference is <= -1e9 ("logzero class"): new must be <= -1e9 as well.
"""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ATOL = 1e-4
RTOL = 2e-6
LZ_CLASS = -1e9
WINDOW_CASES = ["window_full_m3", "window_full_m8_w5", "window_partial_m4"]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def to_np(t):
    if hasattr(t, "materialize"):  # LazyForwardVariables of the lazy-state mode
        t = t.materialize()
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def assert_parity(new, ref, what="", atol=ATOL, rtol=RTOL, ref64=None):
    new, ref = to_np(new).astype(np.float64), to_np(ref).astype(np.float64)
    assert new.shape == ref.shape, f"{what}: shape {new.shape} vs {ref.shape}"
    assert not np.isnan(new).any(), f"{what}: NaN in output"
    lz = ref <= LZ_CLASS
    # also treat as logzero-class anything beyond +1e9 (garbage of finished beams: log_psi - (-1e10))
    big = np.abs(ref) >= -LZ_CLASS
    if lz.any():
        assert (new[lz] <= LZ_CLASS).all(), f"{what}: {int((new[lz] > LZ_CLASS).sum())} logzero-class entries are finite"
    hi = big & ~lz
    if hi.any():
        assert (new[hi] >= -LZ_CLASS).all(), f"{what}: +1e9-class entries differ"
    fin = ~big
    err = np.abs(new[fin] - ref[fin])
    tol = atol + rtol * np.abs(ref[fin])
    bad = err > tol
    if bad.any() and ref64 is not None:
        # adjudicate with the reference's own fp64 run: accept where new is at least as close to fp64 as 2x the
        # reference's fp32 rounding noise
        r64 = to_np(ref64).astype(np.float64)[fin]
        noise = np.abs(ref[fin] - r64)
        bad = bad & (np.abs(new[fin] - r64) > np.maximum(tol, 2 * noise))
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} finite entries out of tolerance, max err "
                           f"{err.max():.3e} at ref={ref[fin][err.argmax()]:.6g}")
    return float(err.max()) if err.size else 0.0


def f64_errors(new, ref64):
    """|new - fp64 run| on the finite class of the fp64 run (zero elsewhere); the logzero / +1e9 classes are checked."""
    a, r = to_np(new).astype(np.float64), to_np(ref64).astype(np.float64)
    assert a.shape == r.shape, f"shape {a.shape} vs {r.shape}"
    assert not np.isnan(a).any(), "NaN in output"
    assert (a[r <= LZ_CLASS] <= LZ_CLASS).all(), "logzero class not preserved against the fp64 run"
    assert (a[r >= -LZ_CLASS] >= -LZ_CLASS).all(), "+1e9 class not preserved against the fp64 run"
    return np.abs(a - r) * (np.abs(r) < -LZ_CLASS)


def assert_no_further_from_fp64(new, yardstick, ref64, what="", floor=ATOL, factor=2.0):
    """Two fp32 evaluation orders of the same recursion (e.g. the sequential and the time-parallel state selection) are
    adjudicated by an fp64 run of the reference algorithm: `new` may be at most max(floor, factor x the worst error of
    `yardstick`) away from fp64.  Returns (worst error of new, worst error of the yardstick)."""
    e_new, e_yard = f64_errors(new, ref64), f64_errors(yardstick, ref64)
    bound = max(floor, factor * float(e_yard.max()) if e_yard.size else 0.0)
    assert (e_new.max() if e_new.size else 0.0) <= bound, (f"{what}: {e_new.max():.3e} from fp64, the yardstick implementation is "
                                                           f"{e_yard.max():.3e} from it (bound {bound:.3e})")
    return float(e_new.max()) if e_new.size else 0.0, float(e_yard.max()) if e_yard.size else 0.0


def select_mode(mode=-1):
    """Query (-1) or set (0 sequential, 1 time-parallel) the lazy state-selection kernel of the library; returns the previous mode."""
    from huggingface_asr_b200 import _lib

    return _lib.lib().ctcps_set_select_pscan(mode)


class Backend:
    """What a replay needs from an implementation."""
    device = "cpu"

    def make_scorer(self, x_logp, lens, blank, eos, margin=0):
        raise NotImplementedError

    def make_processor(self, logits, lens, pad, eos, margin, w, W, space=-1, trick=False, trick_w=1.0):
        raise NotImplementedError

    def t(self, a, dtype=None):
        x = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None:
            x = x.to(dtype)
        return x.to(self.device)


def replay_steps(be: Backend, name: str, blank=3, eos=1):
    """Replay a steps_* golden file: scorer-level and processor-level, every step."""
    g = load(name)
    W = int(g["W"])
    w = float(g["ctc_weight"])
    logits, lens = be.t(g["logits"]), be.t(g["lens"])
    trick = dict(space=int(g["space_token_id"]), trick=bool(g["apply_eos_space_trick"]),
                 trick_w=float(g["eos_space_trick_weight"])) if "space_token_id" in g else {}
    proc = be.make_processor(logits.clone(), lens.clone(), blank, eos, 0, w, W, **trick)
    scorer = be.make_scorer(torch.log_softmax(logits, -1), lens.clone(), blank, eos, 0)
    state = None
    worst = {}
    for n in range(int(g["n_steps"])):
        ids = be.t(g[f"input_ids_{n}"])
        sel = None
        if state is not None:
            sel = scorer.index_select_state(state, ids[:, -1].reshape(-1, W))
            worst["sel_r"] = max(worst.get("sel_r", 0), assert_parity(sel[0], g[f"sel_r_{n}"], f"{name} step {n} sel_r",
                                                                     ref64=g.get(f"sel_r_{n}_f64")))
            assert tuple(sel[1].shape) == (ids.shape[0], logits.shape[-1])
            assert_parity(sel[1][:, 0], g[f"sel_s_{n}"], f"{name} step {n} sel_s", ref64=g.get(f"sel_s_{n}_f64"))
        ts, state = scorer(ids, sel)
        assert tuple(state[0].shape) == (logits.shape[1], 2, ids.shape[0], logits.shape[-1])
        # token and joint scores: the north star's plain 1e-4 absolute (no relative term)
        worst["ts"] = max(worst.get("ts", 0), assert_parity(ts, g[f"token_scores_{n}"], f"{name} step {n} token_scores", rtol=0,
                                                           ref64=g.get(f"token_scores_{n}_f64")))
        assert_parity(state[1], g[f"log_psi_{n}"], f"{name} step {n} log_psi", ref64=g.get(f"log_psi_{n}_f64"))
        if f"r_{n}" in g:
            worst["r"] = max(worst.get("r", 0), assert_parity(state[0], g[f"r_{n}"], f"{name} step {n} r",
                                                             ref64=g.get(f"r_{n}_f64")))
        att = be.t(g[f"att_{n}"])
        out = proc(ids, att)
        worst["out"] = max(worst.get("out", 0), assert_parity(out, g[f"out_{n}"], f"{name} step {n} processor out", rtol=0,
                                                             ref64=g.get(f"out_{n}_f64")))
        # the in-place scores[:, pad] = logzero must reach the caller's tensor (ctc_scorer.py:325)
        assert_parity(att, g[f"att_after_{n}"], f"{name} step {n} att in-place", atol=0, rtol=0)
    return worst


def replay_window(be: Backend, name: str, blank=3, eos=1):
    """Replay a window_* golden file: the scorer with margin > 0 and attention weights (ctc_scorer.py:127-136), full
    vocabulary or scoring_ids, states selected with general ids.  Every tensor of every step, and the (f_min, f_max) the state
    carries."""
    g = load(name)
    W, margin, S = int(g["W"]), int(g["margin"]), int(g["S"])
    logits, lens = be.t(g["logits"]), be.t(g["lens"])
    scorer = be.make_scorer(torch.log_softmax(logits, -1), lens.clone(), blank, eos, margin)
    state = None
    worst = {}
    for n in range(int(g["n_steps"])):
        y = be.t(g[f"y_{n}"])
        sids = be.t(g[f"sids_{n}"]) if S > 0 else None
        ts, st = scorer(y, state, scoring_ids=sids, att_w=be.t(g[f"att_w_{n}"]))
        assert (int(st[2]), int(st[3])) == tuple(int(v) for v in g[f"f_{n}"]), f"{name} step {n}: f_min / f_max"
        worst["ts"] = max(worst.get("ts", 0), assert_parity(ts, g[f"ts_{n}"], f"{name} step {n} token_scores", rtol=0,
                                                           ref64=g.get(f"ts_{n}_f64")))
        worst["r"] = max(worst.get("r", 0), assert_parity(st[0], g[f"r_{n}"], f"{name} step {n} r", ref64=g.get(f"r_{n}_f64")))
        assert_parity(st[1], g[f"log_psi_{n}"], f"{name} step {n} log_psi", ref64=g.get(f"log_psi_{n}_f64"))
        state = scorer.index_select_state(st, be.t(g[f"best_{n}"]))
        worst["sel_r"] = max(worst.get("sel_r", 0), assert_parity(state[0], g[f"sel_r_{n}"], f"{name} step {n} sel_r",
                                                                 ref64=g.get(f"sel_r_{n}_f64")))
        assert_parity(state[1][:, 0], g[f"sel_s_{n}"], f"{name} step {n} sel_s", ref64=g.get(f"sel_s_{n}_f64"))
        assert (int(state[2]), int(state[3])) == tuple(int(v) for v in g[f"f_{n}"])
    return worst


def replay_partial(be: Backend, blank=3, eos=1):
    g = load("partial_scoring")
    W = int(g["W"])
    logits, lens = be.t(g["logits"]), be.t(g["lens"])
    scorer = be.make_scorer(torch.log_softmax(logits, -1), lens, blank, eos, 0)
    BW = logits.shape[0] * W
    ts0, st0 = scorer([[0]] * BW, None, scoring_ids=be.t(g["ids0"]))
    assert_parity(ts0, g["ts0"], "partial ts0", ref64=g["ts0_f64"])
    assert_parity(st0[0], g["r0"], "partial r0", ref64=g["r0_f64"])
    assert_parity(st0[1], g["log_psi0"], "partial log_psi0", ref64=g["log_psi0_f64"])
    assert (to_np(st0[4]) == g["idmap0"]).all()
    sel = scorer.index_select_state(st0, be.t(g["best"]))
    assert_parity(sel[0], g["sel_r"], "partial sel_r", ref64=g["sel_r_f64"])
    assert_parity(sel[1][:, 0], g["sel_s"], "partial sel_s", ref64=g["sel_s_f64"])
    ts1, st1 = scorer(be.t(g["y1"]), sel, scoring_ids=be.t(g["ids1"]))
    assert_parity(ts1, g["ts1"], "partial ts1", ref64=g["ts1_f64"])
    assert_parity(st1[0], g["r1"], "partial r1", ref64=g["r1_f64"])
    assert_parity(st1[1], g["log_psi1"], "partial log_psi1", ref64=g["log_psi1_f64"])
    assert (to_np(st1[4]) == g["idmap1"]).all()


def replay_select_general(be: Backend, blank=3, eos=1):
    g = load("select_general")
    W = int(g["W"])
    logits, lens = be.t(g["logits"]), be.t(g["lens"])
    scorer = be.make_scorer(torch.log_softmax(logits, -1), lens, blank, eos, 0)
    state = (be.t(g["r"]), be.t(g["log_psi"]), 0, 0, None)
    sel = scorer.index_select_state(state, be.t(g["best"]))
    # a gather: bit-exact
    assert_parity(sel[0], g["sel_r"], "select_general r", atol=0, rtol=0)
    assert_parity(sel[1][:, 0], g["sel_s"], "select_general s", atol=0, rtol=0)
    assert tuple(sel[1].shape) == (logits.shape[0] * W, logits.shape[2])


def replay_edges(be: Backend, blank=3, eos=1):
    g = load("edges")
    # (1) start == T-1, == T, > T
    logits, lens = be.t(g["e1_logits"]), be.t(g["e1_lens"])
    W = int(g["e1_W"])
    T = logits.shape[1]
    scorer = be.make_scorer(torch.log_softmax(logits, -1), lens, blank, eos, 0)
    sel_s = be.t(g["e1_sel_s"])
    sel = (be.t(g["e1_sel_r"]), sel_s.view(-1, 1).expand(-1, logits.shape[-1]), 0, 0)
    for L in (T - 1, T, T + 1, T + 3):
        y = [[0] + [5] * L, [0] + [6] * L]
        ts, st = scorer(y, sel)
        assert_parity(ts, g[f"e1_ts_L{L}"], f"edge start L={L} ts")
        assert_parity(st[1], g[f"e1_log_psi_L{L}"], f"edge start L={L} log_psi")
        assert_parity(st[0], g[f"e1_r_L{L}"], f"edge start L={L} r")
    # (2) zero-length and length-1 utterances through the processor
    logits, lens = be.t(g["e2_logits"]), be.t(g["e2_lens"])
    W = int(g["e2_W"])
    proc = be.make_processor(logits.clone(), lens, blank, eos, 0, 0.3, W)
    ids = torch.zeros((logits.shape[0] * W, 1), dtype=torch.long, device=be.device)
    out0 = proc(ids, be.t(g["e2_att0"]))
    assert_parity(out0, g["e2_out0"], "edge len0 out0")
    out1 = proc(be.t(g["e2_ids2"]), be.t(g["e2_att1"]))
    assert_parity(out1, g["e2_out1"], "edge len0 out1")
    assert_parity(proc.ctc_states[1], g["e2_log_psi1"], "edge len0 log_psi1")
    assert_parity(proc.ctc_states[0], g["e2_r1"], "edge len0 r1")
    # (3) token_scores == 0 -> logzero
    logits, lens = be.t(g["e3_logits"]), be.t(g["e3_lens"])
    scorer = be.make_scorer(torch.log_softmax(logits, -1), lens, blank, eos, 0)
    ts, _ = scorer([[0]], (be.t(g["e3_r_prev"]), be.t(g["e3_s_prev"]), 0, 0))
    ref = g["e3_ts"]
    new = to_np(ts)
    # exact zeros depend on bit-identical log_psi; require the class only where the reference hit the hack
    hack = ref <= LZ_CLASS
    assert ((new[hack] <= LZ_CLASS) | (np.abs(new[hack]) <= ATOL)).all()
    assert (np.abs(new[~hack] - ref[~hack]) <= ATOL).all()


def replay_extend(be: Backend, blank=3, eos=1):
    """Streaming helpers extend_prob / extend_state (ctc_scorer.py:209-256) on the reference's single-utterance use."""
    g = load("extend")
    T1, tok = int(g["T1"]), int(g["tok"])
    x_full = be.t(g["x_full"])
    scorer = be.make_scorer(x_full[:, :T1].clone().contiguous(), torch.tensor([T1]), blank, eos, 0)
    V = x_full.shape[-1]
    sel_s = be.t(g["sel_s"])
    scorer.extend_prob(x_full.clone())
    ext = scorer.extend_state((be.t(g["sel_r"]).squeeze(2), sel_s.view(-1, 1).expand(-1, V), 0, 0))
    assert_parity(ext[0], g["ext_r"], "extend_state r_prev")
    assert tuple(ext[0].shape) == tuple(g["ext_r"].shape)
    if hasattr(scorer, "x") and not callable(getattr(scorer, "x")):
        assert_parity(scorer.x, g["x_after"], "x after extend_prob", atol=2e-6, rtol=0)
    ts2, st2 = scorer([[0, tok]], (ext[0].unsqueeze(2), ext[1], 0, 0))
    assert_parity(ts2, g["ts2"], "token_scores after extend")
    assert_parity(st2[0], g["r2"], "r after extend")
    assert_parity(st2[1], g["log_psi2"], "log_psi after extend")


def replay_decode(be: Backend, blank=3, eos=1, bos=0):
    from huggingface_asr_b200.beam_search import joint_beam_search
    from huggingface_asr_b200.synthetic import make_attention_scores

    g = load("decode_1best")
    for i in range(3):
        logits, lens = be.t(g[f"d{i}_logits"]), be.t(g[f"d{i}_lens"])
        W, seed = int(g[f"d{i}_W"]), int(g[f"d{i}_seed"])
        B, T, V = logits.shape
        proc = be.make_processor(logits.clone(), lens, blank, eos, 0, 0.3, W)
        out = joint_beam_search(proc, lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5).to(be.device),
                                B, W, V, bos, eos, blank, max_length=int(g[f"d{i}_max_length"]), device=be.device)
        assert out.steps == int(g[f"d{i}_steps"]), f"decode {i}: steps {out.steps} vs {int(g[f'd{i}_steps'])}"
        assert (to_np(out.lengths) == g[f"d{i}_len"]).all(), f"decode {i}: lengths differ"
        assert (to_np(out.sequences) == g[f"d{i}_seq"]).all(), f"decode {i}: 1-best token sequences differ"
        assert np.abs(to_np(out.scores) - g[f"d{i}_score"]).max() <= 1e-4


PREBEAM_CASES = ["prebeam_w4_s6", "prebeam_w3_s40_v129", "prebeam_w5_s8_tokens_only", "prebeam_w1_s5_flat"]


class _PreBeamChecker:
    """Wraps a processor in pre-beam mode inside the shared harness and compares, at every step, what it was given and
    what it returned with the golden trace of the reference scorer under the same policy (make_golden.case_prebeam)."""

    def __init__(self, proc, g, name):
        self.proc, self.g, self.name, self.n = proc, g, name, 0
        self.worst = {}
        self.use_beam_idx = getattr(proc, "use_beam_idx", False)
        self._sel = None
        scorer = proc.ctc_prefix_scorer
        inner = scorer.index_select_state

        def recording_select(state, best_ids, *a, **k):
            self._sel = inner(state, best_ids, *a, **k)
            return self._sel

        scorer.index_select_state = recording_select

    def set_beam_idx(self, beam_idx):
        self.proc.set_beam_idx(beam_idx)

    def _max(self, key, v):
        self.worst[key] = max(self.worst.get(key, 0.0), v)

    def __call__(self, input_ids, scores):
        g, n, name = self.g, self.n, self.name
        assert (to_np(input_ids) == g[f"input_ids_{n}"]).all(), f"{name} step {n}: the decode took a different path"
        self._sel = None
        out = self.proc(input_ids, scores)
        self._max("out", assert_parity(out, g[f"out_{n}"], f"{name} step {n} processor out", ref64=g.get(f"out_{n}_f64")))
        st = self.proc.ctc_states
        self._max("log_psi", assert_parity(st[1], g[f"log_psi_{n}"], f"{name} step {n} log_psi", ref64=g.get(f"log_psi_{n}_f64")))
        assert tuple(st[0].shape) == (self.proc.ctc_prefix_scorer.input_length, 2, input_ids.shape[0], int(g["S"]))
        if n > 0:
            assert self._sel is not None, f"{name} step {n}: no state selection happened"
            self._max("sel_r", assert_parity(self._sel[0], g[f"sel_r_{n}"], f"{name} step {n} sel_r", ref64=g.get(f"sel_r_{n}_f64")))
            self._max("sel_s", assert_parity(self._sel[1][:, 0], g[f"sel_s_{n}"], f"{name} step {n} sel_s",
                                             ref64=g.get(f"sel_s_{n}_f64")))
        self.n += 1
        return out


def replay_prebeam(make_processor, device, name, blank=3, eos=1, bos=0, teacher_forced=False):
    """make_processor(logits, lens, w, W, S, use_beam_idx) -> a processor in pre-beam mode on `device`.

    teacher_forced: feed the golden input_ids of every step instead of letting the harness choose the beams.  Needed for
    the token-only selection case on another device than the one that made the golden file: there a beam whose token
    was not scored for hypothesis 0 gets s_prev = logzero, all its candidates score 3e9 (one fp32 ulp = 256) and tie
    exactly, and torch.topk breaks exact ties differently on CPU and CUDA."""
    from huggingface_asr_b200.beam_search import joint_beam_search
    from huggingface_asr_b200.synthetic import make_attention_scores

    g = load(name)
    W, S, seed = int(g["W"]), int(g["S"]), int(g["seed"])
    logits = torch.from_numpy(g["logits"]).to(device)
    lens = torch.from_numpy(g["lens"]).to(device)
    B, T, V = logits.shape
    proc = make_processor(logits.clone(), lens, float(g["ctc_weight"]), W, S, bool(g["use_beam_idx"]))
    chk = _PreBeamChecker(proc, g, name)
    if teacher_forced:
        assert not bool(g["use_beam_idx"])
        for n in range(int(g["steps"])):
            chk(torch.from_numpy(g[f"input_ids_{n}"]).to(device), make_attention_scores(B * W, V, n, seed=seed, scale=0.5).to(device))
        return chk.worst
    out = joint_beam_search(chk, lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5).to(device),
                            B, W, V, bos, eos, blank, max_length=int(g["max_length"]), device=device)
    assert out.steps == int(g["steps"])
    assert (to_np(out.sequences) == g["seq"]).all(), f"{name}: 1-best token sequences differ"
    assert (to_np(out.lengths) == g["len"]).all()
    assert np.abs(to_np(out.scores) - g["score"]).max() <= 1e-4
    return chk.worst
