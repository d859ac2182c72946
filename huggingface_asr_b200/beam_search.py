"""Joint CTC/attention beam search around a pluggable LogitsProcessor.

The reference never implements beam search itself: `JointCTCAttentionEncoderDecoder.generate`
(src/models/ctc_encoder_plus_autoregressive_decoder.py:450-482) hands the processor built in
`_get_logits_processor` (:360-404) to transformers 4.39.3 `GenerationMixin.beam_search`, which
cannot run here (SURVEY.md probe table).  This module restates that loop's contract with the
processor -- and nothing else of HF -- as device-agnostic tensor code, so that ONE loop drives the
reference scorer (CPU), the oracle (CPU) and the sm_100a scorer (GPU) on identical inputs:

  * rows are utterance-major (B*W), beam 0 starts at score 0 and the others at -1e9;
  * the processor receives (input_ids (BW,L), log-probs (BW,V)) and returns (BW,V);
  * candidates = top 2W of (B, W*V) after adding the running beam scores;
  * an eos candidate ranked inside the top W is finalised with score / len**length_penalty,
    the first W non-eos candidates continue;
  * rows of finished utterances are fed pad tokens (pad is also the CTC blank);
  * an utterance is done when W hypotheses are finalised and the worst of them beats the best
    score still attainable (early_stopping=False semantics), or at max_length.

Unlike HF there is no per-candidate Python loop and no .item() per utterance: the finished-
hypothesis pool is a (B, W) tensor merged by top-k, and the only host sync is the `all done`
test once per step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import torch


def _finalize(input_ids, beam_scores, pool_scores, pool_seqs, pool_lens, done, B, W, max_length, pad, length_penalty, steps,
              num_return_sequences=1):
    """BeamSearchScorer.finalize: utterances still running contribute their W open beams; the best hypothesis wins.  With
    num_return_sequences = R > 1 the R best of (finished pool, open beams) are returned as well, best first -- what
    do_generate asks of generate() (general_utils.py:197: num_return_sequences, sequences_scores)."""
    dev = input_ids.device
    NEG = float("-inf")
    L = input_ids.shape[1]
    R = int(num_return_sequences)
    if R < 1 or R > W:
        raise ValueError(f"num_return_sequences must be in [1, num_beams = {W}], got {R}")
    open_scores = torch.where(done.view(B, 1), torch.full_like(beam_scores, NEG), beam_scores / float(L - 1) ** length_penalty)
    open_seqs = torch.nn.functional.pad(input_ids.reshape(B, W, L)[:, :, 1:], (0, max_length - (L - 1)), value=pad)
    all_scores = torch.cat([pool_scores, open_scores], dim=1)
    all_seqs = torch.cat([pool_seqs, open_seqs], dim=1)
    all_lens = torch.cat([pool_lens, torch.full((B, W), L - 1, dtype=torch.long, device=dev)], dim=1)
    best = all_scores.argmax(dim=1)
    ar = torch.arange(B, device=dev)
    out = BeamSearchOutput(all_seqs[ar, best], all_lens[ar, best], all_scores[ar, best], steps)
    if R > 1:
        # stable descending order: equal scores keep pool-before-open, lower slot first (argmax above picks the same first entry)
        order = torch.sort(all_scores, dim=1, descending=True, stable=True).indices[:, :R]
        out.nbest_scores = torch.gather(all_scores, 1, order)
        out.nbest_lengths = torch.gather(all_lens, 1, order)
        out.nbest_sequences = torch.gather(all_seqs, 1, order.unsqueeze(-1).expand(B, R, all_seqs.shape[-1]))
    return out


@dataclass
class BeamSearchOutput:
    sequences: torch.Tensor  # (B, max_len) int64, pad-filled, without bos
    lengths: torch.Tensor    # (B,) int64
    scores: torch.Tensor     # (B,) fp32, length-normalised
    steps: int               # processor calls made
    nbest_sequences: torch.Tensor | None = None  # (B, R, max_len) when num_return_sequences = R > 1, best first
    nbest_lengths: torch.Tensor | None = None    # (B, R)
    nbest_scores: torch.Tensor | None = None     # (B, R); -inf where an utterance has fewer than R hypotheses


@torch.no_grad()
def joint_beam_search(processor: Callable, decoder_log_probs: Callable[[torch.Tensor, int], torch.Tensor], batch: int,
                      num_beams: int, vocab: int, bos: int, eos: int, pad: int, max_length: int = 512,
                      length_penalty: float = 1.0, device: torch.device | str = "cpu",
                      sync_every: int = 1, num_return_sequences: int = 1) -> BeamSearchOutput:
    """Run the decode loop; `decoder_log_probs(input_ids, step)` -> (B*W, V) log-probs (a fresh tensor)."""
    B, W, V = batch, num_beams, vocab
    dev = torch.device(device)
    input_ids = torch.full((B * W, 1), bos, dtype=torch.long, device=dev)
    beam_scores = torch.zeros(B, W, dtype=torch.float32, device=dev)
    beam_scores[:, 1:] = -1e9
    NEG = float("-inf")
    pool_scores = torch.full((B, W), NEG, dtype=torch.float32, device=dev)
    pool_seqs = torch.full((B, W, max_length), pad, dtype=torch.long, device=dev)
    pool_lens = torch.zeros(B, W, dtype=torch.long, device=dev)
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    rank = torch.arange(2 * W, device=dev)
    row_base = (torch.arange(B, device=dev) * W).view(B, 1)
    set_beam_idx = getattr(processor, "set_beam_idx", None)
    steps = 0

    while True:
        L = input_ids.shape[1]  # prefix length incl. bos == number of tokens a finalised hyp scores on
        log_probs = decoder_log_probs(input_ids, steps)
        proc = processor(input_ids, log_probs)
        steps += 1
        cand = (proc + beam_scores.view(-1, 1)).view(B, W * V)
        top_scores, top_idx = cand.topk(2 * W, dim=1, largest=True, sorted=True)
        src = torch.div(top_idx, V, rounding_mode="floor")
        tok = top_idx - src * V
        is_eos = tok == eos

        # finalise eos candidates ranked inside the top W (not for utterances already done)
        fin = is_eos & (rank < W) & ~done.view(B, 1)
        fin_scores = torch.where(fin, top_scores / float(L) ** length_penalty, torch.full_like(top_scores, NEG))
        prefix = input_ids.view(B, W, L)[:, :, 1:]  # drop bos
        cand_seqs = torch.gather(prefix, 1, src.unsqueeze(-1).expand(B, 2 * W, L - 1)) if L > 1 else prefix.new_zeros(B, 2 * W, 0)
        all_scores = torch.cat([pool_scores, fin_scores], dim=1)  # (B, 3W)
        keep_scores, keep = all_scores.topk(W, dim=1)
        all_seqs = torch.cat([pool_seqs, torch.nn.functional.pad(cand_seqs, (0, max_length - (L - 1)), value=pad)], dim=1)
        all_lens = torch.cat([pool_lens, torch.full((B, 2 * W), L - 1, dtype=torch.long, device=dev)], dim=1)
        pool_seqs = torch.gather(all_seqs, 1, keep.unsqueeze(-1).expand(B, W, max_length))
        pool_lens = torch.gather(all_lens, 1, keep)
        pool_scores = keep_scores

        # the first W non-eos candidates continue
        order = torch.where(is_eos, rank + 2 * W, rank).argsort(dim=1)[:, :W]
        next_scores = torch.gather(top_scores, 1, order)
        next_tok = torch.gather(tok, 1, order)
        next_src = torch.gather(src, 1, order)

        # done test of the utterance (BeamHypotheses.is_done, early_stopping=False)
        full = (pool_scores > NEG).all(dim=1)
        attainable = top_scores[:, 0] / float(L) ** length_penalty
        done = done | (full & (pool_scores.min(dim=1).values >= attainable))

        dmask = done.view(B, 1)
        next_scores = torch.where(dmask, torch.zeros_like(next_scores), next_scores)
        next_tok = torch.where(dmask, torch.full_like(next_tok, pad), next_tok)
        next_src = torch.where(dmask, torch.zeros_like(next_src), next_src)
        beam_idx = (next_src + row_base).view(-1)
        input_ids = torch.cat([input_ids.index_select(0, beam_idx), next_tok.view(-1, 1)], dim=1)
        beam_scores = next_scores
        if set_beam_idx is not None:  # what HF hands to the model's _reorder_cache; the reference's processor never sees it
            set_beam_idx(beam_idx)

        if input_ids.shape[1] >= max_length:
            break
        if steps % sync_every == 0 and bool(done.all()):
            break

    return _finalize(input_ids, beam_scores, pool_scores, pool_seqs, pool_lens, done, B, W, max_length, pad, length_penalty, steps,
                     num_return_sequences)


_RING = 8
_decode_serial = [0]


def _next_tag_base() -> int:
    """Serial number of a decode, shifted above the step bits: the beam-step kernels publish ((tag_base + step) << 32 | #done)
    into the pinned ring, so an entry written late by an EARLIER decode (whose ring block the pinned allocator may have
    handed to this one) can never be mistaken for this decode's."""
    _decode_serial[0] = (_decode_serial[0] + 1) % (1 << 14)
    return _decode_serial[0] << 16  # a multiple of the ring size: the slot of a step is still step % ring


_rings: dict = {}


def _ring_for(device):
    """The pinned done-ring of (device, host thread): persistent, so a step kernel that the host ran ahead of (done_check_lag)
    never writes into memory that has been handed to somebody else after the decode returned; stale entries carry the
    serial number of their decode and never match."""
    import threading

    key = (str(device), threading.get_ident())
    ring = _rings.get(key)
    if ring is None:
        ring = torch.full((_RING,), -1, dtype=torch.int64).pin_memory()
        _rings[key] = ring
    return ring


def _wait_ring(ring_np, tag: int, stream, what: str, timeout_s: float = 120.0) -> int:
    """Spin until the ring slot of `tag` carries it; returns the number of finished utterances.  A kernel that faulted or
    never ran would leave the host spinning for ever: the stream is polled, and an idle stream without the entry (or a
    deadline) raises instead."""
    import time

    slot = tag % _RING
    t0 = None
    spins = 0
    while True:
        v = int(ring_np[slot])
        if v >> 32 == tag:
            return v & 0xFFFFFFFF
        spins += 1
        if spins % 4096 == 0:
            if stream.query():  # everything enqueued has finished (or failed: query raises) and the entry is still missing
                v = int(ring_np[slot])
                if v >> 32 == tag:
                    return v & 0xFFFFFFFF
                raise RuntimeError(f"{what}: the beam step of tag {tag} finished without publishing its done count")
            now = time.perf_counter()
            t0 = t0 or now
            if now - t0 > timeout_s:
                raise RuntimeError(f"{what}: no done count from the GPU after {timeout_s:.0f} s (tag {tag})")


@torch.no_grad()
def joint_beam_search_fused(processor: Callable, decoder_log_probs: Callable[[torch.Tensor, int], torch.Tensor], batch: int,
                            num_beams: int, vocab: int, bos: int, eos: int, pad: int, max_length: int = 512,
                            length_penalty: float = 1.0, device: torch.device | str = "cuda", done_check_lag: int = 0,
                            num_return_sequences: int = 1) -> BeamSearchOutput:
    """Same loop as joint_beam_search with the whole beam update of a step in ONE kernel (ctcps_beam_step).

    The processor boundary is unchanged: `processor(input_ids (BW,L), log_probs (BW,V))` is still called once per step --
    except for a processor in pre-beam mode (pre_beam_size > 0, lazy state), whose sparse form `score_candidates` is used:
    the step then ranks the W*S candidates directly (ctcps_beam_step_candidates) and no (BW,V) tensor is ever written.
    The host never synchronises: the kernel publishes (step, #done utterances) into pinned memory and the loop reads the
    entry of `done_check_lag` steps ago (0 = wait for the current step, the exact semantics of the torch harness;
    k > 0 = let the CPU run k steps ahead of the GPU, at the price of up to k extra steps after everything is done --
    extra steps never change the result, finished utterances are frozen).
    """
    from . import _lib

    L_ = _lib.lib()
    B, W, V = batch, num_beams, vocab
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("joint_beam_search_fused needs a CUDA device (there is no CPU path)")
    NEG = float("-inf")
    ids = [torch.full((B * W, max_length), pad, dtype=torch.long, device=dev) for _ in range(2)]
    ids[0][:, 0] = bos
    beam_scores = torch.zeros(B, W, dtype=torch.float32, device=dev)
    beam_scores[:, 1:] = -1e9
    pool_scores = torch.full((B, W), NEG, dtype=torch.float32, device=dev)
    pool_seqs = torch.full((B, W, max_length), pad, dtype=torch.long, device=dev)
    pool_lens = torch.zeros(B, W, dtype=torch.long, device=dev)
    done = torch.zeros(B, dtype=torch.uint8, device=dev)
    import ctypes

    nws = ctypes.c_size_t(0)
    _lib.check(L_.ctcps_beam_step_workspace_bytes(B, W, ctypes.byref(nws)), "ctcps_beam_step_workspace_bytes")
    ws = torch.zeros((nws.value + 15) // 16 * 2, dtype=torch.int64, device=dev)  # zeroed: holds the arrival tickets
    RING = _RING
    ring = _ring_for(dev)
    ring_np = ring.numpy()
    tag_base = _next_tag_base()
    stream = torch.cuda.current_stream(dev)
    prefetch = getattr(processor, "prefetch_state", None)
    wants_best = bool(getattr(processor, "use_beam_idx", False)) and prefetch is not None
    best_ids = torch.empty((B, W), dtype=torch.long, device=dev) if wants_best else None
    sparse = (getattr(processor, "pre_beam_size", 0) >= 2 and not getattr(processor, "materialize_state", True)
              and not getattr(processor, "apply_eos_space_trick", False) and hasattr(processor, "score_candidates"))
    cur, L, steps = 0, 1, 0
    while True:
        input_ids = ids[cur][:, :L]
        log_probs = decoder_log_probs(input_ids, steps)
        common = (beam_scores.data_ptr(), ids[cur].data_ptr(), ids[1 - cur].data_ptr(), max_length, L, B, W, V, eos, pad,
                  float(L) ** length_penalty, pool_scores.data_ptr(), pool_lens.data_ptr(), pool_seqs.data_ptr(), max_length,
                  done.data_ptr(), ws.data_ptr(), ws.numel() * 8, ring.data_ptr(), RING, tag_base + steps,
                  None if best_ids is None else best_ids.data_ptr(), stream.cuda_stream)
        if sparse:
            cand_ids, cand_joint = processor.score_candidates(input_ids, log_probs)
            with torch.cuda.device(dev):
                _lib.check(L_.ctcps_beam_step_candidates(cand_joint.data_ptr(), cand_ids.data_ptr(), int(cand_ids.shape[1]), *common),
                           "ctcps_beam_step_candidates")
        else:
            proc = processor(input_ids, log_probs)
            if not proc.is_contiguous():
                proc = proc.contiguous()
            with torch.cuda.device(dev):
                _lib.check(L_.ctcps_beam_step(proc.data_ptr(), *common), "ctcps_beam_step")
        cur ^= 1
        L += 1
        steps += 1
        if L >= max_length:
            break
        if prefetch is not None:  # state selection of the next step overlaps the decoder's forward pass
            if wants_best:
                prefetch(ids[cur][:, :L], best_ids)
            else:
                prefetch(ids[cur][:, :L])
        look = steps - 1 - done_check_lag
        if look >= 0:
            if done_check_lag == 0:
                stream.synchronize()
            # the entry of an older step: wait (briefly) until the GPU has published it
            if _wait_ring(ring_np, tag_base + look, stream, "joint_beam_search_fused") == B:
                break
    return _finalize(ids[cur][:, :L], beam_scores, pool_scores, pool_seqs, pool_lens, done.bool(), B, W, max_length, pad,
                     length_penalty, steps, num_return_sequences)


@torch.no_grad()
def joint_beam_search_native(processor, decoder_log_probs: Callable[[torch.Tensor, int], torch.Tensor], batch: int, num_beams: int,
                             vocab: int, bos: int, eos: int, pad: int, max_length: int = 512, length_penalty: float = 1.0,
                             device: torch.device | str = "cuda", done_check_lag: int = 1, score_timing: list | None = None,
                             fuse_topk: bool = True, num_return_sequences: int = 1) -> BeamSearchOutput:
    """joint_beam_search_fused with ONE host call per decode step (ctcps_decode_step): [top-S candidates,] prefix scoring +
    joint combine, beam step and -- on a side stream, under the next decoder forward pass -- the lazy state selection.
    Same kernels and results as the fused loop; the 5-10 ctypes / torch calls it makes per step cost more host time
    than the ~150-500 us of GPU work they enqueue.

    `processor` is a CTCRescorerLogitsProcessor in lazy-state mode (full vocabulary or pre-beam); the loop uses its
    posteriors, weights and policy flags and keeps the CTC state itself (processor.ctc_states is not touched).
    fuse_topk (full vocabulary only): rank the candidates inside the scoring kernel's epilogue (per-tile top-2W lists merged
    by the beam step) instead of writing the (BW,V) joint scores and ranking them in a second kernel; same results bit for bit.
    score_timing: a list that receives one (begin, end) CUDA-event pair per step, recorded around the step's scoring call;
    resolve_score_timing(list) turns them into milliseconds after the caller has synchronised (bench.py's roofline)."""
    import ctypes

    from . import _lib

    L_ = _lib.lib()
    B, W, V = batch, num_beams, vocab
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("joint_beam_search_native needs a CUDA device (there is no CPU path)")
    sc = processor.ctc_prefix_scorer
    if processor.materialize_state or processor.apply_eos_space_trick:
        raise RuntimeError("the native decode loop runs the lazy-state scorer without the eos/space trick; use joint_beam_search_fused")
    if sc.batch != B or sc.odim != V or processor.num_beams != W:
        raise ValueError("processor and loop disagree on batch / vocabulary / beams")
    S = int(processor.pre_beam_size)
    if S == 1:
        raise ValueError("pre_beam_size = 1 leaves fewer than 2W candidates: use 0 or >= 2")
    T, BW = sc.input_length, B * W
    NEG = float("-inf")
    f32 = dict(dtype=torch.float32, device=dev)
    i64 = dict(dtype=torch.long, device=dev)
    with torch.cuda.device(dev):
        ids = [torch.full((BW, max_length), pad, **i64) for _ in range(2)]
        ids[0][:, 0] = bos
        last = [torch.full((BW,), bos, **i64), torch.empty((BW,), **i64)]
        beam_scores = torch.zeros(B, W, **f32)
        beam_scores[:, 1:] = -1e9
        pool_scores = torch.full((B, W), NEG, **f32)
        pool_seqs = torch.full((B, W, max_length), pad, **i64)
        pool_lens = torch.zeros(B, W, **i64)
        done = torch.zeros(B, dtype=torch.uint8, device=dev)
        nws = ctypes.c_size_t(0)
        _lib.check(L_.ctcps_beam_step_workspace_bytes(B, W, ctypes.byref(nws)), "ctcps_beam_step_workspace_bytes")
        beam_ws = torch.zeros((nws.value + 15) // 16 * 2, **i64)
        r0 = sc.initial_state(W)
        r_sel = [torch.empty((T, 2, BW), **f32) for _ in range(2)]
        s_sel = [torch.empty((BW,), **f32) for _ in range(2)]
        best_ids = torch.empty((B, W), **i64)
        ws = sc._workspace(W, 0)
        keep = [ids, last, beam_scores, pool_scores, pool_seqs, pool_lens, done, beam_ws, r0, r_sel, s_sel, best_ids, ws]
        sess = _lib.DecodeSession()
        sess.B, sess.W, sess.T, sess.V, sess.S = B, W, T, V, S
        sess.blank, sess.eos, sess.pad = sc.blank, eos, pad
        sess.use_beam_idx = int(bool(processor.use_beam_idx))
        w = float(processor.ctc_weight)
        sess.one_minus_w, sess.w, sess.length_penalty = 1.0 - w, w, float(length_penalty)
        sess.ldx, sess.ldt = sc._ldx, sc._ldt
        fused_topk = S == 0 and fuse_topk and V % 4 == 0 and 2 * W <= 64 and W <= 32
        if S > 0:
            sess.x_vt = sc._token_major().data_ptr()
            cand_ids = [torch.empty((BW, S), **i64) for _ in range(2)]
            cand_att = [torch.empty((BW, S), **f32) for _ in range(2)]
            cand_lp = [torch.empty((BW, S), **f32) for _ in range(2)]
            cand_joint = torch.empty((BW, S), **f32)
            keep += [cand_ids, cand_att, cand_lp, cand_joint]
            for k in range(2):
                sess.cand_ids[k], sess.cand_att[k], sess.cand_log_psi[k] = cand_ids[k].data_ptr(), cand_att[k].data_ptr(), cand_lp[k].data_ptr()
            sess.cand_joint = cand_joint.data_ptr()
        else:
            sess.x_logp = sc._frame_major().data_ptr()
            if fused_topk:
                nl, kk = ctypes.c_int(0), ctypes.c_int(0)
                _lib.check(L_.ctcps_topk_lists_shape(B, W, V, ctypes.byref(nl), ctypes.byref(kk)), "ctcps_topk_lists_shape")
                fused_topk = nl.value * kk.value <= 2048
            log_psi = [torch.empty((BW, V), **f32) for _ in range(2)]
            keep.append(log_psi)
            sess.log_psi[0], sess.log_psi[1] = log_psi[0].data_ptr(), log_psi[1].data_ptr()
            if fused_topk:
                # fused scoring + per-tile top-2W: the joint scores never reach memory, only the tiles' candidate lists
                tile_lists = torch.empty((B, nl.value, kk.value, 2), **f32)
                keep.append(tile_lists)
                sess.tile_lists = tile_lists.data_ptr()
            if not fused_topk or max_length - 1 > T:  # the dense step (also taken once a prefix outgrows the utterance, ol > T)
                joint = torch.empty((BW, V), **f32)
                keep.append(joint)
                sess.joint = joint.data_ptr()
        sess.blank_lp, sess.r0 = sc._blank_lp.data_ptr(), r0.data_ptr()
        for k in range(2):
            sess.r_sel[k], sess.s_sel[k], sess.last_ids[k], sess.ids[k] = r_sel[k].data_ptr(), s_sel[k].data_ptr(), last[k].data_ptr(), ids[k].data_ptr()
        sess.score_ws, sess.score_ws_bytes = ws.data_ptr(), ws.numel()
        sc._ws_gen += 1  # the loop overwrites the scorer's workspace
        sess.beam_scores, sess.ld_ids = beam_scores.data_ptr(), max_length
        sess.pool_scores, sess.pool_lens, sess.pool_seqs, sess.ld_pool = pool_scores.data_ptr(), pool_lens.data_ptr(), pool_seqs.data_ptr(), max_length
        sess.done, sess.beam_ws, sess.beam_ws_bytes = done.data_ptr(), beam_ws.data_ptr(), beam_ws.numel() * 8
        RING = _RING
        ring = _ring_for(dev)
        ring_np = ring.numpy()
        tag_base = _next_tag_base()
        sess.done_ring, sess.ring, sess.best_ids, sess.tag_base = ring.data_ptr(), RING, best_ids.data_ptr(), tag_base
        score_lens = sc._score_lens()  # padded frames of short utterances are not streamed
        sess.xlens = None if score_lens is None else score_lens.data_ptr()
        side, ev_a, ev_b = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(L_.ctcps_async_create(ctypes.byref(side), ctypes.byref(ev_a), ctypes.byref(ev_b)), "ctcps_async_create")
        sess.side_stream, sess.ev_step, sess.ev_select = side, ev_a, ev_b
        stream = torch.cuda.current_stream(dev)
        timing_events = []
        sref = ctypes.byref(sess)
        cur, L, steps = 0, 1, 0
        try:
            while True:
                log_probs = decoder_log_probs(ids[cur][:, :L], steps)
                if not log_probs.is_contiguous() or log_probs.dtype != torch.float32:
                    raise ValueError("the decoder must return contiguous float32 log-probs (BW,V)")
                e0 = e1 = None
                if score_timing is not None:
                    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
                    _lib.check(L_.ctcps_event_create(ctypes.byref(e0)), "ctcps_event_create")
                    _lib.check(L_.ctcps_event_create(ctypes.byref(e1)), "ctcps_event_create")
                    timing_events.append((e0, e1))
                _lib.check(L_.ctcps_decode_step(sref, log_probs.data_ptr(), steps, e0, e1, stream.cuda_stream), "ctcps_decode_step")
                cur ^= 1
                L += 1
                steps += 1
                if L >= max_length:
                    break
                look = steps - 1 - done_check_lag
                if look >= 0:
                    if done_check_lag == 0:
                        stream.synchronize()
                    if _wait_ring(ring_np, tag_base + look, stream, "joint_beam_search_native") == B:
                        break
            out = _finalize(ids[cur][:, :L], beam_scores, pool_scores, pool_seqs, pool_lens, done.bool(), B, W, max_length, pad,
                            length_penalty, steps, num_return_sequences)
        finally:
            # the selection of the last step may still run on the side stream and uses buffers this frame owns: order
            # everything the caller enqueues next behind it (no host synchronisation)
            L_.ctcps_decode_finish(sref, stream.cuda_stream)
            if score_timing is not None:
                score_timing.extend(timing_events)  # resolved by the caller with resolve_score_timing() once the work is done
            L_.ctcps_async_destroy(side, ev_a, ev_b)
            del keep
    return out


def resolve_score_timing(pairs) -> list:
    """Milliseconds of the (begin, end) event pairs collected by joint_beam_search_native(score_timing=...); destroys the
    events.  Call after the device has finished the decodes (torch.cuda.synchronize)."""
    import ctypes

    from . import _lib

    L_ = _lib.lib()
    out = []
    ms = ctypes.c_float(0.0)
    for e0, e1 in pairs:
        if L_.ctcps_event_elapsed_ms(e0, e1, ctypes.byref(ms)) == 0:
            out.append(ms.value)
        L_.ctcps_event_destroy(e0)
        L_.ctcps_event_destroy(e1)
    return out
