"""Generate golden vectors by running the UNMODIFIED reference scorer on CPU.

Run in the dev container only (it needs /root/reference):

    python tests/golden/make_golden.py

It imports /root/reference/src/decoding/ctc_scorer.py (never copied), drives it over seeded
synthetic inputs and writes tests/golden/*.npz.  The reference has no tests or golden vectors
of its own, so these files are what pins the oracle and the CUDA path.
Every case stores fp32 outputs of the reference and, as `*_f64`, the same code run in fp64
(the reference honours x.dtype, ctc_scorer.py:35) to adjudicate rounding disputes.

Case kinds
  steps_*   : processor-level replay.  Per decode step: input_ids, attention log-probs, the
              scorer's token_scores / log_psi, the state selected for the next step
              (index_select_state with the processor's own ids, ctc_scorer.py:326-329), the
              processor output, and for the first and last step the full r tensor.
  partial_* : scorer-level, scoring_ids given (ctc_scorer.py:90-97,117-121,155-162,196-202).
  select_*  : index_select_state with general hyp*V+tok ids (ctc_scorer.py:180-207).
  edge_*    : start == T-1, output_length >= T (early return :138-145), all-equal scores
              (token_scores == 0 -> logzero, :176), zero-length-padded utterances.
  decode_*  : 1-best sequences of the shared beam-search harness with the reference processor.
  lm_fusion_*: the reference's LMRescorerLogitsProcessor (shallow_fussion.py, full-prefix LM forward per step) with a small
              random GPT-2 (weights stored in the fixture), over a beam search that reorders hypotheses and a greedy decode.
  window_*  : the reference SCORER with margin > 0 and attention weights (the frame window of ctc_scorer.py:127-136, dead code
              in the reference's processor but part of the scorer's interface): per-step replay, full vocabulary and
              scoring_ids, states selected with general ids.
  prebeam_* : the reference SCORER driven with ESPnet's pre-beam policy (scoring_ids = top-S decoder tokens per
              hypothesis, states selected with source hypothesis * V + token): per-step replay and 1-best decodes.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from decoding.ctc_scorer import CTCPrefixScoreTH, CTCRescorerLogitsProcessor  # noqa: E402  (the reference)

from huggingface_asr_b200.beam_search import joint_beam_search  # noqa: E402
from huggingface_asr_b200.synthetic import BLANK, BOS, EOS, make_attention_scores, make_encoder_logits  # noqa: E402

torch.set_num_threads(8)


def replay_steps(logits, lens, W, n_steps, ctc_weight, seed, dtype, trick=None, force_tokens=None):
    """Drive the reference processor for n_steps with a plain top-W beam update; record everything."""
    B, T, V = logits.shape
    logits = logits.to(dtype)
    trick = trick or dict(space_token_id=-1, apply_eos_space_trick=False, eos_space_trick_weight=1.0)
    proc = CTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, ctc_weight, W, **trick)
    # a second, independent scorer instance to record the scorer-level tensors
    scorer = CTCPrefixScoreTH(torch.log_softmax(logits.clone(), -1), lens.clone(), BLANK, EOS, 0)
    x_padded = scorer.x[0].transpose(0, 1).contiguous()  # (B,T,V) log-posteriors after padding
    input_ids = torch.full((B * W, 1), BOS, dtype=torch.long)
    beam_scores = torch.zeros(B, W, dtype=dtype)
    beam_scores[:, 1:] = -1e9
    state = None
    rec = {"x_padded": x_padded.numpy(), "n_steps": n_steps}
    for n in range(n_steps):
        att = make_attention_scores(B * W, V, n, seed=seed).to(dtype)
        rec[f"input_ids_{n}"] = input_ids.numpy().copy()
        rec[f"att_{n}"] = att.numpy().copy()
        if state is not None:
            sel = scorer.index_select_state(state, input_ids[:, -1].reshape(-1, W))
            rec[f"sel_r_{n}"] = sel[0].numpy().copy()
            rec[f"sel_s_{n}"] = sel[1][:, 0].numpy().copy()
            assert bool((sel[1] == sel[1][:, :1]).all())
        else:
            sel = None
        ts, state = scorer(input_ids, sel)
        rec[f"token_scores_{n}"] = ts.numpy().copy()
        rec[f"log_psi_{n}"] = state[1].numpy().copy()
        if n in (0, n_steps - 1):
            rec[f"r_{n}"] = state[0].numpy().copy()
        att_in = att.clone()
        out = proc(input_ids, att_in)
        rec[f"out_{n}"] = out.numpy().copy()
        rec[f"att_after_{n}"] = att_in.numpy().copy()
        # plain beam update (no eos bookkeeping needed for a replay)
        cand = (out + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        if force_tokens is not None and n in force_tokens:
            tok = torch.full_like(tok, force_tokens[n])
        beam_idx = (src + (torch.arange(B) * W).view(B, 1)).view(-1)
        input_ids = torch.cat([input_ids[beam_idx], tok.view(-1, 1)], dim=1)
        beam_scores = top
    return rec


def with_f64(fn, *a, **k):
    r32 = fn(*a, dtype=torch.float32, **k)
    r64 = fn(*a, dtype=torch.float64, **k)
    for key, v in r64.items():
        if isinstance(v, np.ndarray) and v.dtype == np.float64 and (key.startswith(("token_scores", "log_psi", "sel_", "out_", "r_"))):
            r32[key + "_f64"] = v
    return r32


def save(name, rec):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def case_steps():
    specs = [
        # name, B, W, T, V, kind, ragged, steps, seed
        ("steps_peaky_w3", 2, 3, 40, 48, "peaky", False, 6, 11),
        ("steps_peaky_ragged_w10", 3, 10, 48, 64, "peaky", True, 7, 12),
        ("steps_flat_w1", 2, 1, 33, 40, "flat", True, 5, 13),
        ("steps_flat_w20", 1, 20, 64, 37, "flat", False, 5, 14),   # V not a multiple of 4
        ("steps_peaky_w5_v129", 2, 5, 30, 129, "peaky", True, 5, 15),
    ]
    for name, B, W, T, V, kind, ragged, steps, seed in specs:
        logits, lens, _ = make_encoder_logits(B, T, V, kind, ragged, seed=seed)
        rec = with_f64(replay_steps, logits, lens, W, steps, 0.3, seed)
        rec.update(logits=logits.numpy(), lens=lens.numpy(), W=W, ctc_weight=0.3)
        save(name, rec)
    # finished beams: force pad (= blank) as last token, then continue (garbage-class values)
    logits, lens, _ = make_encoder_logits(2, 24, 32, "peaky", False, seed=16)
    rec = with_f64(replay_steps, logits, lens, 3, 5, 0.3, 16, force_tokens={2: BLANK})
    rec.update(logits=logits.numpy(), lens=lens.numpy(), W=3, ctc_weight=0.3)
    save("steps_forced_pad", rec)
    # eos/space trick on (ctc_scorer.py:333-349)
    logits, lens, _ = make_encoder_logits(2, 24, 32, "flat", False, seed=17)
    trick = dict(space_token_id=7, apply_eos_space_trick=True, eos_space_trick_weight=0.8)
    rec = with_f64(replay_steps, logits, lens, 4, 4, 0.5, 17, trick=trick)
    rec.update(logits=logits.numpy(), lens=lens.numpy(), W=4, ctc_weight=0.5, **trick)
    save("steps_trick", rec)


def case_partial_and_select():
    B, W, T, V, S = 2, 3, 28, 40, 6
    for dtype, suf in ((torch.float32, ""), (torch.float64, "_f64")):
        logits, lens, _ = make_encoder_logits(B, T, V, "peaky", True, seed=21)
        x = torch.log_softmax(logits.to(dtype), -1)
        scorer = CTCPrefixScoreTH(x.clone(), lens, BLANK, EOS, 0)
        g = torch.Generator().manual_seed(5)
        rec = {} if suf == "" else rec  # noqa: F821
        y = [[BOS]] * (B * W)
        ids0 = torch.stack([torch.randperm(V, generator=g)[:S] for _ in range(B * W)])
        ts0, st0 = scorer(y, None, scoring_ids=ids0)
        # general ESPnet-style best ids: hyp*V + tok, tok drawn from the scored set of that hyp (and one outside)
        best = torch.zeros(B, W, dtype=torch.long)
        for b in range(B):
            for w in range(W):
                hyp = int(torch.randint(0, W, (1,), generator=g))
                tok = int(ids0[b * W + hyp][int(torch.randint(0, S, (1,), generator=g))])
                best[b, w] = hyp * V + tok
        best[0, 0] = 1 * V + int([v for v in range(V) if v not in ids0[1].tolist()][0])  # unscored token -> idx 0
        sel = scorer.index_select_state(st0, best)
        y1 = [[BOS, int(best.view(-1)[i] % V)] for i in range(B * W)]
        ids1 = torch.stack([torch.randperm(V, generator=g)[:S] for _ in range(B * W)])
        ids1[:, 0] = torch.tensor([yy[-1] for yy in y1])  # make the last label a scored candidate for every hyp
        ids1[1, 0] = (ids1[1, 1] + 1) % V if (ids1[1, 1] + 1) % V not in ids1[1].tolist() else ids1[1, 0]
        ts1, st1 = scorer(y1, sel, scoring_ids=ids1)
        if suf == "":
            rec.update(logits=logits.numpy(), lens=lens.numpy(), W=W, S=S, ids0=ids0.numpy(), ids1=ids1.numpy(),
                       best=best.numpy(), y1=np.asarray(y1))
        rec.update({f"ts0{suf}": ts0.numpy(), f"r0{suf}": st0[0].numpy(), f"log_psi0{suf}": st0[1].numpy(),
                    f"idmap0{suf}": st0[4].numpy(), f"sel_r{suf}": sel[0].numpy(), f"sel_s{suf}": sel[1][:, 0].numpy(),
                    f"ts1{suf}": ts1.numpy(), f"r1{suf}": st1[0].numpy(), f"log_psi1{suf}": st1[1].numpy(),
                    f"idmap1{suf}": st1[4].numpy()})
    save("partial_scoring", rec)

    # full-vocab select with general ids
    B, W, T, V = 3, 4, 20, 24
    logits, lens, _ = make_encoder_logits(B, T, V, "flat", True, seed=22)
    scorer = CTCPrefixScoreTH(torch.log_softmax(logits, -1), lens, BLANK, EOS, 0)
    ts, st = scorer([[BOS]] * (B * W), None)
    g = torch.Generator().manual_seed(6)
    best = torch.randint(0, W * V, (B, W), generator=g)
    sel = scorer.index_select_state(st, best)
    save("select_general", dict(logits=logits.numpy(), lens=lens.numpy(), W=W, best=best.numpy(), r=st[0].numpy(),
                                log_psi=st[1].numpy(), sel_r=sel[0].numpy(), sel_s=sel[1][:, 0].numpy()))


def case_edges():
    rec = {}
    # (1) long prefixes on a short utterance: start == T-1, start == T, start > T (early return)
    B, W, T, V = 1, 2, 6, 16
    logits, lens, _ = make_encoder_logits(B, T, V, "flat", False, seed=31)
    scorer = CTCPrefixScoreTH(torch.log_softmax(logits, -1), lens, BLANK, EOS, 0)
    ts, st = scorer([[BOS]] * (B * W), None)
    rec.update(e1_logits=logits.numpy(), e1_lens=lens.numpy(), e1_W=W)
    sel = scorer.index_select_state(st, torch.tensor([[5, 6]]))
    rec.update(e1_sel_r=sel[0].numpy(), e1_sel_s=sel[1][:, 0].numpy())
    for L in (T - 1, T, T + 1, T + 3):  # output_length = L
        y = [[BOS] + [5] * L, [BOS] + [6] * L]
        ts, st = scorer(y, sel)
        rec[f"e1_ts_L{L}"] = ts.numpy()
        rec[f"e1_log_psi_L{L}"] = st[1].numpy()
        rec[f"e1_r_L{L}"] = st[0].numpy()
    # (2) utterance fully padded (len 0) next to a normal one, and len 1
    B, W, T, V = 3, 2, 10, 16
    logits, _, _ = make_encoder_logits(B, T, V, "flat", False, seed=32)
    lens = torch.tensor([0, 1, T])
    proc = CTCRescorerLogitsProcessor(logits.clone(), lens, BLANK, EOS, 0, 0.3, W, -1, False, 1.0)
    ids = torch.full((B * W, 1), BOS, dtype=torch.long)
    att = make_attention_scores(B * W, V, 0, seed=32)
    out0 = proc(ids, att.clone())
    ids2 = torch.cat([ids, torch.tensor([[5], [6], [5], [6], [5], [6]])], 1)
    out1 = proc(ids2, make_attention_scores(B * W, V, 1, seed=32))
    rec.update(e2_logits=logits.numpy(), e2_lens=lens.numpy(), e2_W=W, e2_att0=att.numpy(), e2_out0=out0.numpy(),
               e2_ids2=ids2.numpy(), e2_att1=make_attention_scores(B * W, V, 1, seed=32).numpy(), e2_out1=out1.numpy(),
               e2_log_psi1=proc.ctc_states[1].numpy(), e2_r1=proc.ctc_states[0].numpy())
    # (3) token_scores == 0 -> logzero: s_prev equal to log_psi
    B, W, T, V = 1, 1, 8, 12
    logits, lens, _ = make_encoder_logits(B, T, V, "flat", False, seed=33)
    scorer = CTCPrefixScoreTH(torch.log_softmax(logits, -1), lens, BLANK, EOS, 0)
    ts, st = scorer([[BOS]], None)
    r_prev = torch.full((T, 2, 1), -1e10)
    r_prev[:, 1] = torch.cumsum(scorer.x[0, :, :, BLANK], 0)
    ts2, st2 = scorer([[BOS]], (r_prev, st[1].clone(), 0, 0))
    rec.update(e3_logits=logits.numpy(), e3_lens=lens.numpy(), e3_r_prev=r_prev.numpy(), e3_s_prev=st[1].numpy(),
               e3_ts=ts2.numpy())
    save("edges", rec)


def case_decode():
    rec = {}
    for i, (B, W, T, V, kind, ragged) in enumerate([(3, 4, 40, 64, "peaky", True), (2, 10, 56, 96, "peaky", False),
                                                    (2, 3, 24, 32, "flat", True)]):
        logits, lens, transcripts = make_encoder_logits(B, T, V, kind, ragged, seed=40 + i)
        proc = CTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0)
        out = joint_beam_search(proc, lambda ids, n, BW=B * W, V=V, s=40 + i: make_attention_scores(BW, V, n, seed=s, scale=0.5),
                                B, W, V, BOS, EOS, BLANK, max_length=24)
        rec.update({f"d{i}_logits": logits.numpy(), f"d{i}_lens": lens.numpy(), f"d{i}_W": W, f"d{i}_seq": out.sequences.numpy(),
                    f"d{i}_len": out.lengths.numpy(), f"d{i}_score": out.scores.numpy(), f"d{i}_steps": out.steps,
                    f"d{i}_seed": 40 + i, f"d{i}_max_length": 24})
        print("decode", i, "steps", out.steps, "lens", out.lengths.tolist(), "ref transcripts", [len(t) for t in transcripts])
    save("decode_1best", rec)


def case_extend():
    """Streaming helpers extend_prob / extend_state (ctc_scorer.py:209-256), single utterance as in ESPnet's streaming use."""
    T1, T2, V, W = 12, 20, 24, 1
    logits, _, _ = make_encoder_logits(1, T2, V, "flat", False, seed=51)
    x_full = torch.log_softmax(logits, -1)
    scorer = CTCPrefixScoreTH(x_full[:, :T1].clone(), torch.tensor([T1]), BLANK, EOS, 0)
    ts, st = scorer([[BOS]], None)
    tok = 7
    sel = scorer.index_select_state(st, torch.tensor([[tok]]))
    scorer.extend_prob(x_full.clone())
    ext = scorer.extend_state((sel[0].squeeze(2), sel[1], sel[2], sel[3]))
    ts2, st2 = scorer([[BOS, tok]], (ext[0].unsqueeze(2), ext[1], 0, 0))
    save("extend", dict(x_full=x_full.numpy(), T1=T1, tok=tok, sel_r=sel[0].numpy(), sel_s=sel[1][:, 0].numpy(),
                        ext_r=ext[0].numpy(), x_after=scorer.x.numpy(), ts2=ts2.numpy(), r2=st2[0].numpy(), log_psi2=st2[1].numpy()))


class ReferencePreBeamProcessor:
    """The reference processor's arithmetic (ctc_scorer.py:324-332) around the UNMODIFIED reference scorer, with the two
    things ESPnet's beam search does and the reference's processor does not: scoring_ids = top-S tokens of the decoder
    scores (after scores[:, pad] = logzero), and index_select_state ids = source hypothesis * V + token when the loop
    reports its beam indices (set_beam_idx).  use_beam_idx=False keeps the reference's token-only ids (:326-329)."""

    def __init__(self, logits, lens, ctc_weight, W, S, use_beam_idx=True):
        self.scorer = CTCPrefixScoreTH(torch.log_softmax(logits, -1), lens, BLANK, EOS, 0)
        self.w, self.W, self.S, self.use_beam_idx = ctc_weight, W, S, use_beam_idx
        self.states, self._best = None, None
        self.trace = []

    def set_beam_idx(self, beam_idx):
        if self.use_beam_idx:
            self._best = (beam_idx.view(-1, self.W) % self.W) * self.scorer.odim

    def __call__(self, input_ids, scores):
        scores[:, BLANK] = self.scorer.logzero
        sel = None
        if self.states is not None:
            best = input_ids[:, -1].reshape(-1, self.W)
            if self.use_beam_idx and self._best is not None:
                best, self._best = best + self._best, None
            sel = self.scorer.index_select_state(self.states, best)
        ids = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, : self.S].contiguous()
        ctc, self.states = self.scorer(input_ids, sel, scoring_ids=ids)
        out = (1 - self.w) * scores + self.w * ctc
        self.trace.append(dict(ids=ids.numpy().copy(), ctc=ctc.numpy().copy(), out=out.numpy().copy(),
                               log_psi=self.states[1].numpy().copy(),
                               sel_r=None if sel is None else sel[0].numpy().copy(),
                               sel_s=None if sel is None else sel[1][:, 0].numpy().copy()))
        return out


def monotonic_attention(n_bh, T, step, n_steps, seed, dtype):
    """(n_bh, T) attention weights: a soft peak that walks through the utterance with the decode step, a little different
    for every hypothesis."""
    g = torch.Generator().manual_seed(seed * 131 + step)
    centre = (step + 0.5) / n_steps * (T - 1) + torch.randn(n_bh, generator=g) * 1.5
    t = torch.arange(T, dtype=torch.float64)
    w = torch.softmax(-0.5 * ((t.view(1, -1) - centre.view(-1, 1).double()) / 2.0) ** 2, dim=-1)
    return w.to(dtype)


def replay_window(logits, lens, W, margin, n_steps, S, seed, dtype):
    """Drive the reference scorer (margin > 0) with attention weights for n_steps; beam update = top-W of the token scores
    plus noise over the (W, V) candidates of every utterance, general hyp*V+tok ids into index_select_state."""
    B, T, V = logits.shape
    x = torch.log_softmax(logits.to(dtype), -1)
    scorer = CTCPrefixScoreTH(x.clone(), lens.clone(), BLANK, EOS, margin)
    g = torch.Generator().manual_seed(seed)
    y = torch.full((B * W, 1), BOS, dtype=torch.long)
    state = None
    rec = {"n_steps": n_steps}
    for n in range(n_steps):
        att_w = monotonic_attention(B * W, T, n, n_steps, seed, dtype)
        sids = None
        if S > 0:
            sids = torch.stack([torch.randperm(V, generator=g)[:S] for _ in range(B * W)])
            if n > 0:
                sids[:, 0] = y[:, -1]  # the last label is a candidate of every hypothesis but the first
                sids[0, 0] = (sids[0, 1] + 1) % V if (sids[0, 1] + 1) % V not in sids[0].tolist() else sids[0, 0]
        rec[f"y_{n}"] = y.numpy().copy()
        rec[f"att_w_{n}"] = att_w.numpy().copy()
        if sids is not None:
            rec[f"sids_{n}"] = sids.numpy().copy()
        ts, st = scorer([row.tolist() for row in y], state, scoring_ids=sids, att_w=att_w)
        rec[f"ts_{n}"] = ts.numpy().copy()
        rec[f"r_{n}"] = st[0].numpy().copy()
        rec[f"log_psi_{n}"] = st[1].numpy().copy()
        rec[f"f_{n}"] = np.asarray([st[2], st[3]], dtype=np.int64)
        noise = torch.randn(B * W, V, generator=g).to(dtype) * 2.0
        cand = torch.where(ts > -1e9, ts + noise, torch.full_like(ts, -1e9))
        if n == 0:
            cand.view(B, W, V)[:, 1:] = -1e9  # all hypotheses of the first step are the same prefix
        best = cand.view(B, W * V).topk(W, dim=1).indices  # hyp*V + tok
        rec[f"best_{n}"] = best.numpy().copy()
        state = scorer.index_select_state(st, best)
        rec[f"sel_r_{n}"] = state[0].numpy().copy()
        rec[f"sel_s_{n}"] = state[1][:, 0].numpy().copy()
        src = best // V + (torch.arange(B) * W).view(B, 1)
        y = torch.cat([y[src.view(-1)], (best % V).view(-1, 1)], dim=1)
    return rec


def case_window():
    specs = [
        # name, B, W, T, V, kind, ragged, margin, steps, S, seed
        ("window_full_m3", 2, 3, 40, 48, "peaky", True, 3, 6, 0, 41),
        ("window_full_m8_w5", 1, 5, 70, 37, "flat", False, 8, 7, 0, 42),      # V % 4 != 0, windows of several 8-frame chunks
        ("window_partial_m4", 2, 4, 36, 40, "peaky", True, 4, 5, 7, 43),
    ]
    for name, B, W, T, V, kind, ragged, margin, steps, S, seed in specs:
        logits, lens, _ = make_encoder_logits(B, T, V, kind, ragged, seed=seed)
        r32 = replay_window(logits, lens, W, margin, steps, S, seed, torch.float32)
        r64 = replay_window(logits, lens, W, margin, steps, S, seed, torch.float64)
        for key, v in r64.items():
            if isinstance(v, np.ndarray) and v.dtype == np.float64 and key.startswith(("ts_", "log_psi_", "sel_", "r_")):
                r32[key + "_f64"] = v
        # the fp64 run must take the same beam path for its tensors to adjudicate the fp32 ones
        same = all((r32[f"best_{n}"] == r64[f"best_{n}"]).all() and (r32[f"f_{n}"] == r64[f"f_{n}"]).all() for n in range(steps))
        r32.update(logits=logits.numpy(), lens=lens.numpy(), W=W, margin=margin, S=S, f64_same_path=same)
        windows = [tuple(int(v) for v in r32[f"f_{n}"]) for n in range(steps)]
        print(name, "f_min/f_max per step:", windows, "fp64 same path:", same)
        save(name, r32)


def case_prebeam():
    # (1) per-step replay under the shared harness (records what the processor saw and returned at every step)
    specs = [
        # name, B, W, T, V, S, kind, ragged, use_beam_idx, seed, max_length
        ("prebeam_w4_s6", 2, 4, 40, 48, 6, "peaky", True, True, 61, 14),
        ("prebeam_w3_s40_v129", 2, 3, 30, 129, 40, "peaky", False, True, 62, 10),   # S > 32: two-list top-k
        ("prebeam_w5_s8_tokens_only", 2, 5, 36, 64, 8, "peaky", True, False, 63, 12),  # the reference's hyp-0 selection
        ("prebeam_w1_s5_flat", 3, 1, 25, 33, 5, "flat", True, True, 64, 10),
    ]
    for name, B, W, T, V, S, kind, ragged, ubi, seed, max_length in specs:
        logits, lens, _ = make_encoder_logits(B, T, V, kind, ragged, seed=seed)
        rec = dict(logits=logits.numpy(), lens=lens.numpy(), W=W, S=S, ctc_weight=0.3, use_beam_idx=ubi, seed=seed,
                   max_length=max_length)
        for dtype, suf in ((torch.float32, ""), (torch.float64, "_f64")):
            proc = ReferencePreBeamProcessor(logits.to(dtype), lens.clone(), 0.3, W, S, ubi)
            seen = []

            def dec(ids, n, BW=B * W, V=V, s=seed, dtype=dtype, seen=seen):
                seen.append(ids.numpy().copy())
                return make_attention_scores(BW, V, n, seed=s, scale=0.5).to(dtype)

            out = joint_beam_search(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=max_length)
            if suf == "":
                rec.update(seq=out.sequences.numpy(), len=out.lengths.numpy(), score=out.scores.numpy(), steps=out.steps)
                for n, ids in enumerate(seen):
                    rec[f"input_ids_{n}"] = ids
            for n, tr in enumerate(proc.trace):
                for k, v in tr.items():
                    if v is not None and (suf == "" or k != "ids"):
                        rec[f"{k}_{n}{suf}"] = v
        print(name, "steps", rec["steps"], "lens", rec["len"].tolist())
        save(name, rec)


def tiny_gpt2(vocab, seed):
    """A small random GPT-2 LM; its weights are stored in the fixture, so the test does not depend on torch's initialiser."""
    from transformers import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(seed)
    cfg = GPT2Config(vocab_size=vocab, n_positions=64, n_embd=32, n_layer=2, n_head=4, bos_token_id=BOS, eos_token_id=EOS,
                     resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    lm = GPT2LMHeadModel(cfg).eval()
    with torch.no_grad():  # the default init gives a nearly uniform LM; make it opinionated
        for p_ in lm.parameters():
            p_.mul_(6.0)
    return lm


def case_lm_fusion():
    """LMRescorerLogitsProcessor (src/decoding/shallow_fussion.py:5-58), the reference's full-prefix version, replayed over a
    beam search that reorders and duplicates hypotheses (B=3, W=4) and over a greedy decode (W=1)."""
    from decoding.shallow_fussion import LMRescorerLogitsProcessor  # the reference

    for name, B, W, n_steps in (("lm_fusion_w4", 3, 4, 9), ("lm_fusion_w1", 2, 1, 6)):
        V = 48
        lm = tiny_gpt2(V, seed=11)
        proc = LMRescorerLogitsProcessor(0.5, lm, torch.device("cpu"))
        rec = {"n_steps": n_steps, "B": B, "W": W, "V": V, "lm_weight": 0.5}
        for k, v in lm.state_dict().items():
            rec["lm." + k] = v.numpy().copy()
        input_ids = torch.full((B * W, 1), BOS, dtype=torch.long)
        beam_scores = torch.zeros(B, W)
        beam_scores[:, 1:] = -1e9
        for n in range(n_steps):
            att = make_attention_scores(B * W, V, n, seed=21, scale=0.5)
            with torch.no_grad():
                out = proc(input_ids, att.clone())
            rec[f"input_ids_{n}"] = input_ids.numpy().copy()
            rec[f"att_{n}"] = att.numpy().copy()
            rec[f"out_{n}"] = out.numpy().copy()
            cand = (out + beam_scores.view(-1, 1)).view(B, W * V)
            top, idx = cand.topk(W, dim=1)
            src, tok = idx // V, idx % V
            beam_idx = (src + (torch.arange(B) * W).view(B, 1)).view(-1)
            rec[f"beam_idx_{n}"] = beam_idx.numpy().copy()
            input_ids = torch.cat([input_ids[beam_idx], tok.view(-1, 1)], dim=1)
            beam_scores = top
        save(name, rec)


if __name__ == "__main__":
    if "--prebeam-only" in sys.argv:
        case_prebeam()
        sys.exit(0)
    if "--lm-only" in sys.argv:
        case_lm_fusion()
        sys.exit(0)
    if "--window-only" in sys.argv:
        case_window()
        sys.exit(0)
    case_steps()
    case_partial_and_select()
    case_edges()
    case_decode()
    case_extend()
    case_window()
    case_prebeam()
    case_lm_fusion()
