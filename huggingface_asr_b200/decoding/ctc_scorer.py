"""B200-native drop-in for BUTSpeechFIT/huggingface_asr `src/decoding/ctc_scorer.py`.

Same module layout, class names, constructor signatures, call signatures, return shapes and
in-place side effects as the reference, so `JointCTCAttentionEncoderDecoder._get_logits_processor`
(src/models/ctc_encoder_plus_autoregressive_decoder.py:382-397, src/reguler/modeling_decred.py:456-475)
can import this module instead of its own:

    CTCPrefixScoreTH            (reference ctc_scorer.py:7)    __call__ :58, index_select_state :180,
                                                               extend_prob :209, extend_state :231
    CTCRescorerLogitsProcessor  (reference ctc_scorer.py:259)  __call__ :324
    LogSoftmaxProcessor         (reference ctc_scorer.py:357)

Everything numerical happens in hand-written sm_100a kernels behind the C ABI of include/ctcps.h
(libctcps_b200.so, loaded with ctypes).  torch supplies device memory and the current stream only.
There is no CPU path: CPU tensors raise.  No call synchronises the host.
"""
from __future__ import annotations

import os

import torch
from transformers import LogitsProcessor

from .. import _lib

LOGZERO = -10000000000.0


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(t: torch.Tensor, name: str, dim: int | None = None) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the sm_100a CTC prefix scorer has no CPU path")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32 (the kernels compute in fp32 log space), got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise ValueError(f"{name} must have {dim} dimensions, got shape {tuple(t.shape)}")


class LazyForwardVariables:
    """Stands in for the forward variables r (T,2,BW,V) of a step scored in lazy-state mode.

    The scores of a step do not depend on r, and index_select_state keeps BW of its BW*V columns, so in lazy mode r is
    never written; index_select_state recomputes the surviving columns from what this object remembers (the inputs of the
    step).  `materialize()` produces the full tensor on demand (runs the materialising kernel), so code that inspects
    `ctc_states[0]` still finds the reference's tensor.
    """

    def __init__(self, scorer, r_prev, last_ids, ol, n_hyps, scoring_ids=None):
        self.scorer, self.r_prev, self.last_ids, self.ol, self.n_hyps = scorer, r_prev, last_ids, ol, n_hyps
        self.scoring_ids = scoring_ids  # (BW,S) when the step was scored on candidates only
        snum = scorer.odim if scoring_ids is None else int(scoring_ids.shape[1])
        self.shape = (scorer.input_length, 2, scorer.batch * n_hyps, snum)
        self.dtype, self.device = torch.float32, scorer.device

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    def materialize(self):
        _, state, _ = self.scorer._launch_score(self.r_prev, None, self.last_ids, self.ol, self.n_hyps, self.scoring_ids, None, 0.0,
                                                False)
        return state[0]

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __getattr__(self, name):  # anything else a tensor can do (.cpu(), .sum(), ...): do it on the materialised tensor
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class _SelectedState(tuple):
    """(r_prev, s_prev, f_min, f_max) of index_select_state, plus the generation of the scorer workspace that the lazy
    select already prepared for the next scoring call (None if it did not)."""
    prepared_gen = None


class CandidateState:
    """State of a step that was scored on `scoring_ids` candidates in lazy mode (pre-beam decoding).  Holds only what the
    next index_select_state needs -- the inputs of the step, the candidate ids and their prefix scores -- and behaves like
    the reference's 5-tuple (r (T,2,BW,S), log_psi (BW,V), 0, 0, scoring_idmap (BW,V)) when indexed: those tensors are built
    on demand (inspection only; the decode path never touches them)."""

    def __init__(self, scorer, r_prev, last_ids, ol, n_hyps, scoring_ids, cand_log_psi):
        self.r = LazyForwardVariables(scorer, r_prev, last_ids, ol, n_hyps, scoring_ids)
        self.scoring_ids, self.cand_log_psi = scoring_ids, cand_log_psi

    def __len__(self):
        return 5

    def __iter__(self):
        return (self[i] for i in range(5))

    def __getitem__(self, i):
        if i == 0:
            return self.r
        if i in (2, 3):
            return 0
        n_bh, V = self.scoring_ids.shape[0], self.r.scorer.odim
        if i == 1:  # log_psi: logzero outside the candidates (:156, :161-162)
            return torch.full((n_bh, V), LOGZERO, dtype=torch.float32, device=self.cand_log_psi.device).scatter_(
                1, self.scoring_ids, self.cand_log_psi)
        if i == 4:  # scoring_idmap (:91-95)
            S = self.scoring_ids.shape[1]
            return torch.full((n_bh, V), -1, dtype=torch.long, device=self.scoring_ids.device).scatter_(
                1, self.scoring_ids, torch.arange(S, device=self.scoring_ids.device).expand(n_bh, S))
        raise IndexError(i)


class CTCPrefixScoreTH(object):
    """Batched CTC prefix scorer (Watanabe et al. Algorithm 2, vectorised over hypotheses), on sm_100a.

    Interface of the reference class (ctc_scorer.py:7-256).  Logical shapes of everything returned match
    the reference; the internal layout differs where the reference's is pure redundancy:
      * the reference materialises self.x = stack([x, blank column broadcast]) (2,T,B,V) (:44-46); here the
        padded log-posteriors stay (B,T,V) plus a (B,T) blank column, and `.x` builds the reference tensor on demand;
      * s_new of index_select_state is a (BW,) vector expanded to (BW,V) (the reference repeats it, :194).
    """

    def __init__(self, x, xlens, blank, eos, margin=0):
        """x: (B,T,V) log-posteriors, padded IN PLACE like the reference (:39-42); xlens: (B,) lengths."""
        _require_cuda_f32(x, "x", 3)
        if not x.is_contiguous():
            raise ValueError("x must be contiguous (B,T,V)")
        self._setup(x, xlens, blank, eos, margin, apply_log_softmax=False)

    @classmethod
    def from_logits(cls, logits, xlens, blank, eos, margin=0, token_major=False):
        """Fused log-softmax + padding (K-a): what CTCRescorerLogitsProcessor.__init__ does with
        F.log_softmax(encoder_logits) (reference :278-284) in one pass, leaving `logits` untouched.
        token_major=True keeps the posteriors as (B,V,ldt) -- a token's time series contiguous -- which is the layout the
        pre-beam (candidate) kernels gather from; the frame-major copy is then rebuilt only if a full-vocabulary call needs it."""
        if isinstance(logits, torch.Tensor) and logits.is_floating_point() and logits.dtype != torch.float32:
            # the reference log-softmaxes whatever dtype the encoder produced (fp16 / bf16 under autocast); the fused
            # K-a writes a new fp32 buffer anyway, so half-precision encoder outputs are upcast instead of rejected
            logits = logits.float()
        _require_cuda_f32(logits, "encoder_logits", 3)
        self = cls.__new__(cls)
        self._setup(logits.contiguous(), xlens, blank, eos, margin, apply_log_softmax=True, token_major=token_major)
        return self

    @classmethod
    def from_hidden_states(cls, hidden, head, xlens, blank, eos, margin=0, token_major=False):
        """the scorer built from the encoder's last hidden states (B,T,d) and its CTC head (ctc_head.CTCHead):
        the head kernel writes the padded log-posteriors and the blank column itself, K-a is not run."""
        self = cls.__new__(cls)
        L = _lib.lib()
        self.logzero = LOGZERO
        self.blank, self.eos, self.margin = int(blank), eos, margin
        self.batch, self.input_length, self.odim = int(hidden.shape[0]), int(hidden.shape[1]), int(head.vocab)
        self.dtype, self.device = torch.float32, hidden.device
        if not 0 <= self.blank < self.odim:
            raise ValueError(f"blank id {blank} is outside the CTC vocabulary of size {self.odim}")
        lens = torch.as_tensor(xlens)
        if lens.numel() != self.batch:
            raise ValueError(f"xlens has {lens.numel()} entries for a batch of {self.batch}")
        self.end_frames = lens - 1
        self._lens = lens.to(device=self.device, dtype=torch.long).contiguous()
        self._ldx = L.ctcps_padded_ld(self.odim)
        self._xt, self._ldt = None, L.ctcps_padded_lt(self.input_length)
        self._x, self._blank_lp = head.log_posteriors(hidden, self._lens, self.blank)
        if token_major:
            self._token_major()
        self._finish_setup(margin)
        return self

    def _setup(self, x, xlens, blank, eos, margin, apply_log_softmax, token_major=False):
        L = _lib.lib()
        self.logzero = LOGZERO
        self.blank = int(blank)
        self.eos = eos
        self.batch, self.input_length, self.odim = (int(s) for s in x.shape)
        self.dtype = x.dtype
        self.device = x.device
        self.margin = margin
        if not 0 <= self.blank < self.odim:
            raise ValueError(f"blank id {blank} is outside the CTC vocabulary of size {self.odim}")
        B, T, V = self.batch, self.input_length, self.odim
        lens = torch.as_tensor(xlens)
        if lens.numel() != B:
            raise ValueError(f"xlens has {lens.numel()} entries for a batch of {B}")
        self.end_frames = lens - 1  # :47
        self._lens = lens.to(device=self.device, dtype=torch.long).contiguous()
        ldx = L.ctcps_padded_ld(V)
        self._ldx = ldx
        self._xt, self._ldt = None, L.ctcps_padded_lt(T)
        with torch.cuda.device(self.device):
            self._blank_lp = torch.empty((B, T), dtype=torch.float32, device=self.device)
            if apply_log_softmax or ldx != V:
                self._x = torch.empty((B, T, ldx), dtype=torch.float32, device=self.device)
            else:
                self._x = x  # V % 64 == 0: the caller's tensor is the storage (padded in place, like the reference)
            if not apply_log_softmax and ldx != V:
                # pad the caller's tensor in place first (reference semantics), then stage the strided copy
                _lib.check(L.ctcps_init(_ptr(x), V, _ptr(self._lens), B, T, V, self.blank, 0, _ptr(x), V, None,
                                        _stream(self.device)), "ctcps_init")
            _lib.check(L.ctcps_init(_ptr(x), V, _ptr(self._lens), B, T, V, self.blank, int(apply_log_softmax),
                                    _ptr(self._x), ldx, _ptr(self._blank_lp), _stream(self.device)), "ctcps_init")
        if token_major:
            with torch.cuda.device(self.device):
                self._xt = torch.empty((B, V, self._ldt), dtype=torch.float32, device=self.device)
                _lib.check(L.ctcps_transpose_vt(_ptr(self._x), ldx, B, T, V, _ptr(self._xt), self._ldt, _stream(self.device)),
                           "ctcps_transpose_vt")
            if self._x is not x:
                self._x = None  # rebuilt on demand by _frame_major()
        self._finish_setup(margin)

    def _finish_setup(self, margin):
        B, T, V = self.batch, self.input_length, self.odim
        self._ws = None
        self._ws_key = None
        self._ws_gen = 0  # bumped by every call that overwrites the workspace
        self._timing = None
        self.lazy_state = False  # True: never materialise r (see LazyForwardVariables)
        self.idx_bh = None
        self.idx_b = torch.arange(B, device=self.device)      # :55
        self.idx_bo = (self.idx_b * V).unsqueeze(1)           # :56
        self.scoring_num = 0
        if margin > 0:
            self.frame_ids = torch.arange(T, dtype=self.dtype, device=self.device)  # :52

    # -- reference attribute, built on demand --------------------------------------------------------
    @property
    def x(self):
        """(2,T,B,V) tensor of the reference (:44-46).  Costs 2x the posteriors; only for inspection."""
        xn = self._frame_major()[:, :, : self.odim].transpose(0, 1)
        xb = self._blank_lp.transpose(0, 1).unsqueeze(2).expand(-1, -1, self.odim)
        return torch.stack([xn, xb])

    def _frame_major(self):
        """The (B,T,ldx) log-posteriors the full-vocabulary kernels stream through TMA; a token-major scorer builds them
        from its (B,V,ldt) copy the first time a full-vocabulary call (or an inspection of .x / r) asks."""
        if self._x is None:
            T, V = self.input_length, self.odim
            self._x = torch.nn.functional.pad(self._xt[:, :, :T].transpose(1, 2), (0, self._ldx - V)).contiguous()
        return self._x

    def _token_major(self):
        if self._xt is None:
            B, T, V = self.batch, self.input_length, self.odim
            with torch.cuda.device(self.device):
                self._xt = torch.empty((B, V, self._ldt), dtype=torch.float32, device=self.device)
                _lib.check(_lib.lib().ctcps_transpose_vt(_ptr(self._x), self._ldx, B, T, V, _ptr(self._xt), self._ldt,
                                                         _stream(self.device)), "ctcps_transpose_vt")
        return self._xt

    def _score_lens(self):
        """(B) int64 lengths for the scoring kernels, or None when they do not describe the current posteriors (extend_prob)."""
        lens = getattr(self, "_lens", None)
        if lens is None or lens.numel() != self.batch or not lens.is_cuda:
            return None
        return lens

    def _workspace(self, W, S):
        key = (self.batch, self.input_length, W, S)
        if self._ws_key != key:
            import ctypes

            n = ctypes.c_size_t(0)
            _lib.check(_lib.lib().ctcps_workspace_bytes(self.batch, self.input_length, self.odim, W, S, ctypes.byref(n)),
                       "ctcps_workspace_bytes")
            self._ws = torch.empty((max(int(n.value), 256),), dtype=torch.uint8, device=self.device)
            self._ws_key = key
        return self._ws

    def initial_state(self, n_hyps):
        """r_prev of the empty prefix, (T,2,B*n_hyps) (reference :74-85)."""
        r0 = torch.empty((self.input_length, 2, self.batch * n_hyps), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctcps_initial_state(_ptr(self._blank_lp), self.batch, self.input_length, n_hyps, 0, _ptr(r0),
                                                      _stream(self.device)), "ctcps_initial_state")
        return r0

    def __call__(self, y, state, scoring_ids=None, att_w=None):
        """Compute CTC prefix scores for next labels (reference :58-178).

        y: prefix label sequences, a (BW,L) LongTensor or a list of sequences (only y[i][-1] and len(y[0]) are used);
        state: None or (r_prev (T,2,BW), s_prev (BW,V) | 0.0, f_min, f_max);  scoring_ids: optional (BW,S).
        Returns (token_scores (BW,V), (r (T,2,BW,S or V), log_psi (BW,V), 0, 0, scoring_idmap | None)).
        """
        ts, new_state, _ = self._score(y, state, scoring_ids, att_w, None, 0.0)
        return ts, new_state

    def _parse_step(self, y, state, scoring_ids, att_w):
        """Shared argument handling of a scoring call: (n_bh, ol, last_ids, W, S, scoring_ids, r_prev, s_prev)."""
        dev = self.device
        B, T = self.batch, self.input_length
        if isinstance(y, torch.Tensor):
            if y.dim() != 2:
                raise ValueError(f"y must be (BW, L), got {tuple(y.shape)}")
            n_bh, ol = int(y.shape[0]), int(y.shape[1]) - 1
            last_ids = y[:, -1].to(device=dev, dtype=torch.long).contiguous()
        else:
            n_bh, ol = len(y), len(y[0]) - 1                       # :68-70
            last_ids = torch.tensor([int(yi[-1]) for yi in y], dtype=torch.long).to(dev)
        if n_bh % B != 0:
            raise ValueError(f"{n_bh} hypotheses are not a multiple of the batch of {B} utterances")
        W = n_bh // B                                            # :71
        S = 0
        if scoring_ids is not None:
            if not scoring_ids.is_cuda:
                raise RuntimeError("scoring_ids must be a CUDA tensor")
            scoring_ids = scoring_ids.to(torch.long).contiguous()
            if scoring_ids.dim() != 2 or scoring_ids.shape[0] != n_bh:
                raise ValueError(f"scoring_ids must be (BW,S) with BW={n_bh}, got {tuple(scoring_ids.shape)}")
            S = int(scoring_ids.shape[1])
        self.scoring_num = S                                     # :72
        if state is None:
            r_prev, s_prev = self.initial_state(W), None
        else:
            r_prev, s_prev = state[0], state[1]
            _require_cuda_f32(r_prev, "state r_prev")
            if tuple(r_prev.shape) != (T, 2, n_bh):
                raise ValueError(f"state r_prev must be {(T, 2, n_bh)}, got {tuple(r_prev.shape)}")
            r_prev = r_prev.contiguous()
        return n_bh, ol, last_ids, W, S, scoring_ids, r_prev, s_prev

    def _window(self, att_w, state, ol, n_bh):
        """The frame window of reference :127-132: (start, end, f_min, f_max), host scalars there too (two .cpu() reads)."""
        T = self.input_length
        if not isinstance(att_w, torch.Tensor) or not att_w.is_cuda:
            raise RuntimeError("att_w must be a CUDA tensor")
        if tuple(att_w.shape) != (n_bh, T):
            raise ValueError(f"att_w must be {(n_bh, T)}, got {tuple(att_w.shape)}")
        f_min_prev, f_max_prev = (0, 1) if state is None else (int(state[2]), int(state[3]))   # :84-85
        f_arg = torch.matmul(att_w.to(self.dtype), self.frame_ids)
        lo_hi = torch.stack([f_arg.min(), f_arg.max()]).cpu()
        f_min = max(int(lo_hi[0]), f_min_prev)
        f_max = max(int(lo_hi[1]), f_max_prev)
        start = min(f_max_prev, max(f_min - self.margin, ol, 1))
        end = min(f_max + self.margin, T)
        return start, end, f_min, f_max

    def _score(self, y, state, scoring_ids, att_w, att_scores, ctc_weight, need_token_scores=True):
        n_bh, ol, last_ids, W, S, scoring_ids, r_prev, s_prev = self._parse_step(y, state, scoring_ids, att_w)
        if att_w is not None and self.margin > 0:
            # windowed call (dead code in the reference's processor, part of the scorer's interface): the materialising kernels
            # take the window; the state they return is the reference's tensor
            window = self._window(att_w, state, ol, n_bh)
            return self._launch_score(r_prev, s_prev, last_ids, ol, W, scoring_ids, att_scores, ctc_weight, False,
                                      need_token_scores, False, window)
        prepared = (state is not None and getattr(state, "prepared_gen", None) == self._ws_gen
                    and self._ws_key == (self.batch, self.input_length, W, S))
        return self._launch_score(r_prev, s_prev, last_ids, ol, W, scoring_ids, att_scores, ctc_weight,
                                  self.lazy_state and scoring_ids is None, need_token_scores, prepared)

    def _score_candidates(self, y, state, scoring_ids, cand_att, ctc_weight, need_token_scores=False):
        """Lazy-state scoring of `scoring_ids` (BW,S) candidates (unique per hypothesis): the (BW,S) prefix scores, token
        scores and joint scores (with cand_att, the decoder scores of the candidates), and a CandidateState from which
        index_select_state recomputes the survivors' forward variables.  No (T,2,BW,S) state, no (BW,V) tensor."""
        L = _lib.lib()
        dev = self.device
        B, T, V = self.batch, self.input_length, self.odim
        n_bh, ol, last_ids, W, S, scoring_ids, r_prev, s_prev = self._parse_step(y, state, scoring_ids, None)
        if S == 0:
            raise ValueError("_score_candidates needs scoring_ids")
        s_vec = None
        if isinstance(s_prev, torch.Tensor):
            _require_cuda_f32(s_prev, "state s_prev")
            s_vec = (s_prev if s_prev.dim() == 1 else s_prev[:, 0]).contiguous()  # the reference's s_prev is a row broadcast (:194)
            if s_vec.numel() != n_bh:
                raise ValueError(f"state s_prev must have {n_bh} rows, got {tuple(s_prev.shape)}")
        elif s_prev is not None and float(s_prev) != 0.0:
            s_vec = torch.full((n_bh,), float(s_prev), dtype=torch.float32, device=dev)
        # the workspace layout does not depend on S: a select of the previous step prepared it for this call
        prepared = (state is not None and getattr(state, "prepared_gen", None) == self._ws_gen
                    and self._ws_key == (B, T, W, 0))
        if cand_att is not None:
            _require_cuda_f32(cand_att, "cand_att", 2)
            if tuple(cand_att.shape) != (n_bh, S) or not cand_att.is_contiguous():
                raise ValueError(f"cand_att must be contiguous {(n_bh, S)}, got {tuple(cand_att.shape)}")
        xt = self._token_major()
        w = float(ctc_weight)
        with torch.cuda.device(dev):
            cand_log_psi = torch.empty((n_bh, S), dtype=torch.float32, device=dev)
            cand_ts = torch.empty((n_bh, S), dtype=torch.float32, device=dev) if need_token_scores else None
            cand_joint = torch.empty((n_bh, S), dtype=torch.float32, device=dev) if cand_att is not None else None
            ws = self._workspace(W, 0)
            self._ws_gen += 1
            timing = self._timing
            if timing is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            _lib.check(L.ctcps_score_candidates(_ptr(xt), self._ldt, _ptr(r_prev), _ptr(s_vec), _ptr(last_ids), ol, B, W, T, V,
                                                self.blank, _ptr(scoring_ids), S, _ptr(cand_att), 1.0 - w, w, _ptr(cand_log_psi),
                                                _ptr(cand_ts), _ptr(cand_joint), _ptr(ws), ws.numel(), int(bool(prepared)),
                                                _stream(dev)), "ctcps_score_candidates")
            if timing is not None:
                ev1.record()
                timing.append((ev0, ev1))
        return cand_log_psi, cand_ts, cand_joint, CandidateState(self, r_prev, last_ids, ol, W, scoring_ids, cand_log_psi), s_vec

    def _launch_score(self, r_prev, s_prev, last_ids, ol, W, scoring_ids, att_scores, ctc_weight, lazy, need_token_scores=True,
                      prepared=False, window=None):
        L = _lib.lib()
        dev = self.device
        B, T, V = self.batch, self.input_length, self.odim
        n_bh = B * W
        S = 0 if scoring_ids is None else int(scoring_ids.shape[1])
        snum = S if S > 0 else V
        s_ptr, s_rs, s_cs = None, 0, 0
        if isinstance(s_prev, torch.Tensor):
            _require_cuda_f32(s_prev, "state s_prev")
            if s_prev.dim() == 1:
                s_prev = s_prev.view(-1, 1).expand(n_bh, V)
            if tuple(s_prev.shape) != (n_bh, V):
                raise ValueError(f"state s_prev must be {(n_bh, V)}, got {tuple(s_prev.shape)}")
            s_ptr, (s_rs, s_cs) = s_prev.data_ptr(), s_prev.stride()
        elif s_prev is not None and float(s_prev) != 0.0:
            s_prev = torch.full((n_bh,), float(s_prev), dtype=torch.float32, device=dev)
            s_ptr, s_rs, s_cs = s_prev.data_ptr(), 1, 0

        joint = None
        if att_scores is not None:
            _require_cuda_f32(att_scores, "scores", 2)
            if tuple(att_scores.shape) != (n_bh, V) or not att_scores.is_contiguous():
                raise ValueError(f"scores must be contiguous {(n_bh, V)}, got {tuple(att_scores.shape)}")
        ldr = L.ctcps_padded_ld(snum)
        with torch.cuda.device(dev):
            log_psi = torch.empty((n_bh, V), dtype=torch.float32, device=dev)
            # the processor only needs the joint scores; skipping token_scores saves a (BW,V) write per step
            skip_ts = not need_token_scores and att_scores is not None and S == 0 and ol <= T
            token_scores = None if skip_ts else torch.empty((n_bh, V), dtype=torch.float32, device=dev)
            if att_scores is not None:
                joint = torch.empty((n_bh, V), dtype=torch.float32, device=dev)
            idmap = torch.empty((n_bh, V), dtype=torch.long, device=dev) if S > 0 else None
            ws = self._workspace(W, S)
            self._ws_gen += 1
            w = float(ctc_weight)
            timing = self._timing
            if timing is not None:  # bench.py: CUDA events around the K-b launches on the launching stream
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            x_fm = self._frame_major()
            if lazy:
                r = LazyForwardVariables(self, r_prev, last_ids, ol, W)
                # the lengths let the kernel leave out the padded frames of short utterances (they add exactly 0)
                _lib.check(L.ctcps_score_lazy_lens(_ptr(x_fm), self._ldx, _ptr(self._blank_lp), _ptr(self._score_lens()), _ptr(r_prev),
                                                   s_ptr, s_rs, s_cs, _ptr(last_ids), ol, B, W, T, V, self.blank, _ptr(att_scores),
                                                   1.0 - w, w, _ptr(log_psi), _ptr(token_scores), _ptr(joint), _ptr(ws), ws.numel(),
                                                   int(bool(prepared)), _stream(dev)), "ctcps_score_lazy_lens")
            else:
                r = torch.empty((T, 2, n_bh, ldr), dtype=torch.float32, device=dev)
                start, end = (max(ol, 1), T) if window is None else window[:2]
                _lib.check(L.ctcps_score_window(_ptr(x_fm), self._ldx, _ptr(self._blank_lp), _ptr(r_prev), s_ptr, s_rs, s_cs,
                                                _ptr(last_ids), ol, start, end, B, W, T, V, self.blank, _ptr(scoring_ids), S,
                                                _ptr(idmap), _ptr(att_scores), 1.0 - w, w, _ptr(r), ldr, _ptr(log_psi),
                                                _ptr(token_scores), _ptr(joint), _ptr(ws), ws.numel(), _stream(dev)),
                           "ctcps_score_window")
                if ldr != snum:
                    r = r[..., :snum]
            if timing is not None:
                ev1.record()
                timing.append((ev0, ev1))
        f_min, f_max = (0, 0) if window is None else window[2:]
        return token_scores, (r, log_psi, f_min, f_max, idmap), joint

    def index_select_state(self, state, best_ids, _out=None):
        """Select CTC states according to best ids (reference :180-207).

        best_ids: (B,W) ids in hyp*V + tok space.  Returns (r_new (T,2,BW), s_new (BW,V) [expanded], f_min, f_max).
        """
        if isinstance(state, CandidateState):
            return self._select_candidates(state, best_ids, _out)
        r, s, f_min, f_max, scoring_idmap = state
        _require_cuda_f32(s, "state log_psi", 2)
        if isinstance(r, LazyForwardVariables):
            n_bh = int(s.shape[0])
            best_ids = best_ids.to(device=self.device, dtype=torch.long).contiguous()
            if best_ids.numel() != n_bh:
                raise ValueError(f"best_ids has {best_ids.numel()} entries for {n_bh} hypotheses")
            T, V = self.input_length, self.odim
            s = s.contiguous()
            with torch.cuda.device(self.device):
                if _out is not None:  # caller-owned buffers (prefetch_state: no allocation on the side stream)
                    r_new, s_vec = _out
                else:
                    r_new = torch.empty((T, 2, n_bh), dtype=torch.float32, device=self.device)
                    s_vec = torch.empty((n_bh,), dtype=torch.float32, device=self.device)
                # the scan also prepares the workspace of the scoring call that follows (same W, full vocabulary)
                ws = self._workspace(r.n_hyps, 0) if self.lazy_state else None
                _lib.check(_lib.lib().ctcps_select_lazy(_ptr(self._frame_major()), self._ldx, _ptr(self._blank_lp), _ptr(r.r_prev),
                                                        _ptr(r.last_ids), r.ol, _ptr(s), _ptr(best_ids), self.batch, r.n_hyps, T, V,
                                                        _ptr(r_new), _ptr(s_vec), _ptr(ws), 0 if ws is None else ws.numel(),
                                                        _stream(self.device)), "ctcps_select_lazy")
            out = _SelectedState((r_new, s_vec.view(-1, 1).expand(n_bh, V), f_min, f_max))
            if ws is not None and r.ol + 1 <= T:
                self._ws_gen += 1
                out.prepared_gen = self._ws_gen
            return out
        _require_cuda_f32(r, "state r", 4)
        T, _, n_bh, snum = (int(v) for v in r.shape)
        V = self.odim
        n_hyps = n_bh // self.batch
        S = 0 if scoring_idmap is None else snum
        if r.stride(3) == 1 and r.stride(1) == n_bh * r.stride(2) and r.stride(0) == 2 * n_bh * r.stride(2):
            ldr = int(r.stride(2))
        else:
            r = r.contiguous()
            ldr = snum
        best_ids = best_ids.to(device=self.device, dtype=torch.long).contiguous()
        if best_ids.numel() != n_bh:
            raise ValueError(f"best_ids has {best_ids.numel()} entries for {n_bh} hypotheses")
        s = s.contiguous()
        with torch.cuda.device(self.device):
            if _out is not None:
                r_new, s_vec = _out
            else:
                r_new = torch.empty((T, 2, n_bh), dtype=torch.float32, device=self.device)
                s_vec = torch.empty((n_bh,), dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib().ctcps_select(_ptr(r), ldr, _ptr(s), _ptr(best_ids), _ptr(scoring_idmap), self.batch, n_hyps,
                                               T, V, S, _ptr(r_new), _ptr(s_vec), _stream(self.device)), "ctcps_select")
        return r_new, s_vec.view(-1, 1).expand(n_bh, V), f_min, f_max

    def _select_candidates(self, state, best_ids, _out=None):
        """index_select_state after a candidate step (reference :196-202 on a state that was never written)."""
        r = state.r
        n_bh, S = (int(v) for v in state.scoring_ids.shape)
        T, V = self.input_length, self.odim
        best_ids = best_ids.to(device=self.device, dtype=torch.long).contiguous()
        if best_ids.numel() != n_bh:
            raise ValueError(f"best_ids has {best_ids.numel()} entries for {n_bh} hypotheses")
        with torch.cuda.device(self.device):
            if _out is not None:
                r_new, s_vec = _out
            else:
                r_new = torch.empty((T, 2, n_bh), dtype=torch.float32, device=self.device)
                s_vec = torch.empty((n_bh,), dtype=torch.float32, device=self.device)
            ws = self._workspace(r.n_hyps, 0)
            _lib.check(_lib.lib().ctcps_select_lazy_candidates(_ptr(self._token_major()), self._ldt, _ptr(self._blank_lp),
                                                               _ptr(r.r_prev), _ptr(r.last_ids), r.ol, _ptr(state.scoring_ids), S,
                                                               _ptr(state.cand_log_psi), _ptr(best_ids), self.batch, r.n_hyps, T, V,
                                                               _ptr(r_new), _ptr(s_vec), _ptr(ws), ws.numel(),
                                                               _stream(self.device)), "ctcps_select_lazy_candidates")
        out = _SelectedState((r_new, s_vec.view(-1, 1).expand(n_bh, V), 0, 0))
        if r.ol + 1 <= T:
            self._ws_gen += 1
            out.prepared_gen = self._ws_gen
        return out

    def extend_prob(self, x):
        """Extend the posteriors with new frames (streaming helper, reference :209-229).  x: (B,T',V) log-posteriors."""
        _require_cuda_f32(x, "x", 3)
        T_old = self.input_length
        if T_old < x.shape[1]:
            B, T_new, V = (int(v) for v in x.shape)
            if B != self.batch or V != self.odim:
                raise ValueError("extend_prob: batch / vocabulary mismatch")
            L = _lib.lib()
            old_x, old_blank = self._frame_major(), self._blank_lp
            self._xt = None
            with torch.cuda.device(self.device):
                self._x = torch.empty((B, T_new, self._ldx), dtype=torch.float32, device=self.device)
                self._blank_lp = torch.empty((B, T_new), dtype=torch.float32, device=self.device)
                # xlens = [T'] in the reference (:218): nothing is padded
                _lib.check(L.ctcps_init(_ptr(x.contiguous()), V, None, B, T_new, V, self.blank, 0, _ptr(self._x), self._ldx,
                                        _ptr(self._blank_lp), _stream(self.device)), "ctcps_init")
            self._x[:, :T_old] = old_x            # frames already seen keep their values (:227)
            self._blank_lp[:, :T_old] = old_blank
            self.input_length = T_new
            self._ldt = L.ctcps_padded_lt(T_new)
            self.end_frames = torch.as_tensor([T_new]) - 1
            self._lens = None  # the new frames are not padding: the scoring kernels must stream all of them
            self._ws_key = None

    def extend_state(self, state):
        """Extend the blank-only chain of a state to the new input length (reference :231-256)."""
        if state is None:
            return state
        r_prev, s_prev, f_min_prev, f_max_prev = state
        _require_cuda_f32(r_prev, "state r_prev")
        squeeze = r_prev.dim() == 2  # the reference's single-hypothesis (T,2) layout
        rp = r_prev.unsqueeze(2) if squeeze else r_prev
        T_old, _, n_bh = (int(v) for v in rp.shape)
        T = self.input_length
        r_new = torch.full((T, 2, n_bh), LOGZERO, dtype=torch.float32, device=self.device)
        start = max(T_old, 1)
        r_new[:T_old] = rp
        if start < T:
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().ctcps_initial_state(_ptr(self._blank_lp), self.batch, T, n_bh // self.batch, start,
                                                          _ptr(r_new), _stream(self.device)), "ctcps_initial_state")
        return (r_new.squeeze(2) if squeeze else r_new, s_prev, f_min_prev, f_max_prev)


class CTCRescorerLogitsProcessor(LogitsProcessor):
    """Joint CTC/attention rescoring processor (reference ctc_scorer.py:259-354), one instance per generate()."""

    def __init__(
        self,
        encoder_logits: torch.FloatTensor,
        encoder_output_lens: torch.LongTensor,
        pad_token_id: int,
        eos_token_id: int,
        ctc_margin: int,
        ctc_weight: float,
        num_beams: int,
        space_token_id: int,
        apply_eos_space_trick: bool,
        eos_space_trick_weight: float,
        debug: bool = False,
        *,
        materialize_state: bool | None = None,
        pre_beam_size: int = 0,
        use_beam_idx: bool | None = None,
    ):
        """Same positional signature as the reference.  Keyword-only extensions (not in the reference):

        materialize_state
            False = lazy state (default): never write r (T,2,BW,V); the W surviving columns per utterance are recomputed
                    at the next step -- same joint scores (ulp-level), bit-identical selected states, ~20x fewer HBM bytes
                    per step.  `ctc_states[0]` is then a LazyForwardVariables that materialises the tensor on demand.
            True  = write the full state every step, exactly the reference's data flow.
            None  = take it from the environment variable CTCPS_MATERIALIZE_STATE (default "0").
        pre_beam_size
            0 (default) = score the full vocabulary, like the reference's processor.  S > 0 = ESPnet's pre-beam: only the
            top-S tokens of the decoder scores of every hypothesis are CTC-scored (the scorer's `scoring_ids` path,
            reference :90-97); every other token gets the reference's logzero-class score.  Changes results (tokens
            outside the candidates can no longer win), hence off by default.
        use_beam_idx
            False = select the state of the next step from token ids only, the reference processor's behaviour (:326-329:
            every beam inherits from hypothesis 0 of its utterance).  True = use the ids handed over by `set_beam_idx` /
            `prefetch_state(best_ids=...)` (source hypothesis * V + token, ESPnet's semantics, reference :180-191).
            None = True when pre_beam_size > 0 (a token is only scored for the hypothesis that proposed it), else False."""
        self._init_common(encoder_logits, encoder_output_lens, pad_token_id, eos_token_id, ctc_margin, ctc_weight, num_beams, space_token_id,
                          apply_eos_space_trick, eos_space_trick_weight, debug, materialize_state=materialize_state,
                          pre_beam_size=pre_beam_size, use_beam_idx=use_beam_idx)

    def _init_common(self, encoder_logits, encoder_output_lens, pad_token_id, eos_token_id, ctc_margin, ctc_weight, num_beams,
                     space_token_id, apply_eos_space_trick, eos_space_trick_weight, debug=False, *, materialize_state=None,
                     pre_beam_size=0, use_beam_idx=None, hidden_states=None, ctc_head=None):
        LogitsProcessor.__init__(self)
        self.pad_token_id = pad_token_id
        self.pre_beam_size = int(pre_beam_size)
        if self.pre_beam_size < 0 or self.pre_beam_size > 64 or self.pre_beam_size == 1:
            # beam search draws 2W candidates out of the W * S scored ones: S = 1 leaves fewer than 2W, and the unscored
            # tokens it would then pick select lane 0 with a logzero prefix score -- every later state would be garbage
            raise ValueError("pre_beam_size must be 0 (full vocabulary) or in [2, 64] (W * S >= 2W candidates are needed)")
        if hidden_states is not None:
            self.ctc_prefix_scorer = CTCPrefixScoreTH.from_hidden_states(hidden_states, ctc_head, encoder_output_lens, pad_token_id,
                                                                        eos_token_id, ctc_margin, token_major=self.pre_beam_size > 0)
        else:
            self.ctc_prefix_scorer = CTCPrefixScoreTH.from_logits(encoder_logits, encoder_output_lens, pad_token_id, eos_token_id,
                                                                  ctc_margin, token_major=self.pre_beam_size > 0)
        if materialize_state is None:
            materialize_state = os.environ.get("CTCPS_MATERIALIZE_STATE", "0") not in ("0", "false", "False", "no")
        self.materialize_state = bool(materialize_state)
        self.ctc_prefix_scorer.lazy_state = not self.materialize_state
        self.use_beam_idx = (self.pre_beam_size > 0) if use_beam_idx is None else bool(use_beam_idx)
        self.ctc_weight = ctc_weight
        self.ctc_states = None
        self._best_ids = None    # (B,W) ids for the next index_select_state, from set_beam_idx()
        self._prev_ids = None    # input_ids of the previous call: parents are recovered from them when nobody reports beam indices
        self._prefetched = None  # (last-token tensor, selected state, event) produced by prefetch_state()
        self._side_stream = None
        self._prefetch_bufs = None
        self._prefetch_turn = 0
        self.num_beams = num_beams
        self.eos_token_id = eos_token_id
        self.apply_eos_space_trick = apply_eos_space_trick
        self.space_token_id = space_token_id
        self.eos_space_trick_weight = eos_space_trick_weight
        self.debug = debug

    @classmethod
    def from_encoder_hidden_states(cls, hidden_states: torch.FloatTensor, ctc_head, encoder_output_lens, *args, **kwargs):
        """take the encoder's last hidden states (B,T,d) and its CTC head (ctc_head.CTCHead) instead of the
        (B,T,V) logits -- a tenth of the bytes -- and compute the logits here (TF32x3 tensor-core GEMM, fp32 accuracy).
        The remaining arguments are those of the constructor after `encoder_output_lens`."""
        self = cls.__new__(cls)
        self._init_common(None, encoder_output_lens, *args, hidden_states=hidden_states, ctc_head=ctc_head, **kwargs)
        return self

    # -- state selection ----------------------------------------------------------------------------------
    def set_beam_idx(self, beam_idx: torch.LongTensor) -> None:
        """Tell the processor which rows of the previous step the current rows continue (HF's `beam_idx`, global row
        indices b*W + source hypothesis; what the model's `_reorder_cache` receives).  Only used with use_beam_idx."""
        if not self.use_beam_idx:
            return
        W = self.num_beams
        self._best_ids = (beam_idx.to(self.ctc_prefix_scorer.device).view(-1, W) % W) * self.ctc_prefix_scorer.odim  # + token, in _select

    def _select_ids(self, input_ids, best_ids=None):
        """ids for index_select_state: ESPnet's hyp*V + tok when the beam indices are known, else token ids only (:326-329)."""
        last = input_ids[:, -1].reshape(-1, self.num_beams)
        if best_ids is not None:
            return best_ids.reshape(-1, self.num_beams)
        if self.use_beam_idx and self._best_ids is not None:
            ids, self._best_ids = self._best_ids + last, None
            return ids
        if self.use_beam_idx and self.ctc_states is not None:
            # Nobody called set_beam_idx() (HF generate() without a KV cache never calls _reorder_cache; a custom loop may
            # not forward beam_idx).  In pre-beam mode a token is only scored for the hypothesis that proposed it, so the
            # reference's token-only selection would read a column that was never computed: recover every row's parent
            # by matching its prefix against the rows of the previous call, or refuse.
            parents = self._parents_by_prefix(input_ids)
            return parents * self.ctc_prefix_scorer.odim + last
        return last

    def _parents_by_prefix(self, input_ids):
        """(B,W) index of the hypothesis of the previous call that every current row extends (first match inside its own
        utterance; rows with equal prefixes carry equal states).  Raises when the previous ids are unknown."""
        prev = self._prev_ids
        W = self.num_beams
        if prev is None or prev.shape[0] != input_ids.shape[0] or prev.shape[1] != input_ids.shape[1] - 1:
            raise RuntimeError("use_beam_idx is set but no beam indices were supplied for this step (call set_beam_idx(beam_idx) from "
                               "the model's _reorder_cache, or prefetch_state(..., best_ids)) and the rows of the previous call are "
                               "not available to recover them from")
        L1 = prev.shape[1]
        cur = input_ids[:, :L1].reshape(-1, W, 1, L1)
        old = prev.reshape(-1, 1, W, L1)
        eq = (cur == old).all(dim=-1)                      # (B, W_cur, W_prev)
        found = eq.any(dim=-1)
        # a row whose prefix matches nothing (cannot happen under beam search) falls back to hypothesis 0 like the reference
        return torch.where(found, eq.to(torch.uint8).argmax(dim=-1), torch.zeros_like(found, dtype=torch.long))

    def _select(self, input_ids):
        if self._prefetched is not None:
            last, selected, done = self._prefetched
            self._prefetched = None
            if last.numel() != input_ids.shape[0]:
                raise RuntimeError("prefetch_state() was called for a different batch than this step")
            torch.cuda.current_stream(self.ctc_prefix_scorer.device).wait_event(done)
            self.ctc_states = selected
        elif self.ctc_states is not None:
            self.ctc_states = self.ctc_prefix_scorer.index_select_state(self.ctc_states, self._select_ids(input_ids))

    def _check_scores(self, scores):
        sc = self.ctc_prefix_scorer
        _require_cuda_f32(scores, "scores", 2)
        if scores.shape[-1] != sc.odim:
            raise ValueError(f"decoder vocabulary ({scores.shape[-1]}) != CTC vocabulary ({sc.odim}): the encoder's extra "
                             "blank_projection column (src/models/encoders/e_branchformer.py:415,456-457) is not supported "
                             "by the reference processor either")
        return scores if scores.is_contiguous() else scores.contiguous()

    def __call__(self, input_ids: torch.LongTensor, scores: torch.FloatTensor) -> torch.FloatTensor:
        sc = self.ctc_prefix_scorer
        work = self._check_scores(scores)
        self._select(input_ids)
        if self.use_beam_idx:
            self._prev_ids = input_ids
        need_ts = self.apply_eos_space_trick or self.debug
        if self.pre_beam_size > 0:
            ctc_scores, next_token_scores = self._call_pre_beam(input_ids, work, need_ts)
        else:
            # scores[:, pad] = logzero (:325), the scorer (:330) and the combine (:332) are one fused launch
            ctc_scores, self.ctc_states, next_token_scores = sc._score(input_ids, self.ctc_states, None, None, work, self.ctc_weight,
                                                                       need_token_scores=need_ts)
        if work is not scores:
            scores[:, self.pad_token_id] = sc.logzero
        if self.apply_eos_space_trick:
            with torch.cuda.device(sc.device):
                _lib.check(_lib.lib().ctcps_eos_space_trick(_ptr(work), _ptr(ctc_scores), _ptr(next_token_scores),
                                                            int(scores.shape[0]), sc.odim, int(self.eos_token_id),
                                                            int(self.space_token_id), float(self.eos_space_trick_weight),
                                                            _stream(sc.device)), "ctcps_eos_space_trick")
        if self.debug:
            self.analyze_predictions(scores, ctc_scores, next_token_scores, input_ids)
        return next_token_scores

    # -- pre-beam (candidate) decoding ----------------------------------------------------------------------
    def _top_candidates(self, work):
        """scores[:, pad] = logzero in place (:325) and the top pre_beam_size (ids, scores) of every row."""
        sc = self.ctc_prefix_scorer
        n_bh, S = int(work.shape[0]), self.pre_beam_size
        with torch.cuda.device(sc.device):
            ids = torch.empty((n_bh, S), dtype=torch.long, device=sc.device)
            cand_att = torch.empty((n_bh, S), dtype=torch.float32, device=sc.device)
            _lib.check(_lib.lib().ctcps_prebeam_topk(_ptr(work), n_bh, sc.odim, sc.blank, S, _ptr(ids), _ptr(cand_att),
                                                     _stream(sc.device)), "ctcps_prebeam_topk")
        return ids, cand_att

    def _call_pre_beam(self, input_ids, work, need_ts):
        """The reference-shaped (BW,V) outputs of a pre-beam step."""
        sc = self.ctc_prefix_scorer
        ids, cand_att = self._top_candidates(work)
        if self.materialize_state:  # the reference's data flow: scorer(..., scoring_ids) writes r (T,2,BW,S) and the idmap
            ctc_scores, self.ctc_states, joint = sc._score(input_ids, self.ctc_states, ids, None, work, self.ctc_weight)
            return ctc_scores, joint
        L = input_ids.shape[1] - 1
        cand_log_psi, cand_ts, cand_joint, self.ctc_states, s_vec = sc._score_candidates(input_ids, self.ctc_states, ids, cand_att,
                                                                                         self.ctc_weight, need_token_scores=need_ts)
        n_bh, V = int(work.shape[0]), sc.odim
        w = float(self.ctc_weight)
        with torch.cuda.device(sc.device):
            joint = torch.empty((n_bh, V), dtype=torch.float32, device=sc.device)
            ctc_scores = torch.empty((n_bh, V), dtype=torch.float32, device=sc.device) if need_ts else None
            _lib.check(_lib.lib().ctcps_candidates_to_dense(_ptr(work), _ptr(s_vec), _ptr(ids), _ptr(cand_log_psi), _ptr(cand_ts),
                                                            _ptr(cand_joint), n_bh, V, self.pre_beam_size, 1.0 - w, w, L,
                                                            sc.input_length, None, _ptr(ctc_scores), _ptr(joint),
                                                            _stream(sc.device)), "ctcps_candidates_to_dense")
        return ctc_scores, joint

    def score_candidates(self, input_ids: torch.LongTensor, scores: torch.FloatTensor):
        """Sparse form of __call__ for a decode loop that owns its beam step (beam_search.joint_beam_search_fused): returns
        (cand_ids (BW,S) int64, cand_joint (BW,S)) -- the joint scores of the pre-beam candidates -- and never builds a (BW,V)
        tensor.  Every other token has the logzero-class joint score __call__ would return and can never be selected.
        Same side effect on `scores` ([:, pad] = logzero) as __call__."""
        if self.pre_beam_size <= 0 or self.materialize_state:
            raise RuntimeError("score_candidates needs pre_beam_size > 0 and lazy state")
        if self.apply_eos_space_trick:
            raise RuntimeError("the eos/space trick needs the dense scores: use __call__")
        work = self._check_scores(scores)
        if work is not scores:
            raise ValueError("scores must be contiguous")
        self._select(input_ids)
        if self.use_beam_idx:
            self._prev_ids = input_ids
        ids, cand_att = self._top_candidates(work)
        _, _, cand_joint, self.ctc_states, _ = self.ctc_prefix_scorer._score_candidates(input_ids, self.ctc_states, ids, cand_att,
                                                                                      self.ctc_weight)
        return ids, cand_joint

    def prefetch_state(self, input_ids: torch.LongTensor, best_ids: torch.LongTensor | None = None) -> None:
        """Optional hook, not in the reference: start the state selection of the NEXT step (reference :326-329) as soon as
        its last tokens are known, on a side stream, so that it overlaps the attention decoder's forward pass.  A decode
        loop that owns its beam search (beam_search.joint_beam_search_fused) calls it right after the beam update; HF's
        loop never does and then __call__ selects the state itself.  The next __call__ must be for these input_ids.
        best_ids (B,W): source hypothesis * V + token of every row (ctcps_beam_step's best_ids_out); used when the
        processor was built with use_beam_idx, else ignored (the reference's token-only selection)."""
        if self.ctc_states is None or self._prefetched is not None:
            return
        sc = self.ctc_prefix_scorer
        main = torch.cuda.current_stream(sc.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(sc.device)
        side = self._side_stream
        # on the main stream, before the hand-over
        last = self._select_ids(input_ids, best_ids if self.use_beam_idx else None).contiguous()
        n_bh = int(input_ids.shape[0])
        if self._prefetch_bufs is None or self._prefetch_bufs[0][0].shape[2] != n_bh:
            # two sets: the state selected for step n is still being read while the one for step n+1 is written
            self._prefetch_bufs = [(torch.empty((sc.input_length, 2, n_bh), dtype=torch.float32, device=sc.device),
                                    torch.empty((n_bh,), dtype=torch.float32, device=sc.device)) for _ in range(2)]
            self._prefetch_turn = 0
        out = self._prefetch_bufs[self._prefetch_turn]
        self._prefetch_turn ^= 1
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            selected = sc.index_select_state(self.ctc_states, last, _out=out)
            done = torch.cuda.Event()
            done.record(side)
        self._prefetched = (last, selected, done)

    @staticmethod
    def analyze_predictions(scores, ctc_scores, next_token_scores, input_ids, k=10, tokenizer=None):
        """Debug dump (reference :294-322 decodes with a hub tokenizer; offline we print ids unless one is given)."""
        dec = (lambda ids: tokenizer.batch_decode(ids)) if tokenizer is not None else (lambda ids: [str(i.tolist()) for i in ids])
        print("PREFIX:")
        for index, prefix in enumerate(dec(input_ids)):
            print(f"HYP {index}:\n{prefix}")
        for name, t in (("ATT_SCORES", scores), ("CTC_SCORES", ctc_scores), ("NEXT_TOKEN_SCORES", next_token_scores)):
            best = t.topk(k=k, dim=1)
            print(f"{name}:")
            for index, (ids, vals) in enumerate(zip(dec(best.indices), best.values)):
                print(f"HYP {index}:\n{ids} {vals}")


class LogSoftmaxProcessor(LogitsProcessor):
    """log_softmax over the vocabulary (reference ctc_scorer.py:357-365); inserted for greedy decoding."""

    def __init__(self):
        super().__init__()

    def __call__(self, input_ids: torch.LongTensor, scores: torch.FloatTensor) -> torch.FloatTensor:
        _require_cuda_f32(scores, "scores", 2)
        src = scores if scores.stride(-1) == 1 else scores.contiguous()
        out = torch.empty(tuple(scores.shape), dtype=torch.float32, device=scores.device)
        with torch.cuda.device(scores.device):
            _lib.check(_lib.lib().ctcps_log_softmax(_ptr(src), int(src.stride(0)), _ptr(out), int(out.stride(0)),
                                                    int(scores.shape[0]), int(scores.shape[1]), _stream(scores.device)),
                       "ctcps_log_softmax")
        return out
