"""GPU: the scoring kernel's fused per-tile top-2W epilogue (ctcps_score_lazy_topk) and the beam step over its candidate lists
(ctcps_beam_step_lists) against the unfused pair they replace in the native decode loop: ctcps_score_lazy (dense joint
tensor) + ctcps_beam_step.  Everything here is exact: same log_psi and keys bit for bit, same ranking, same ties.

Reference lines: ctc_scorer.py:154-176 (log_psi, token scores), :325,:332 (joint combine), :180-207 (what the state
selection reads); the beam step restates HF's beam_search / BeamSearchScorer.process.
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

BLANK, EOS, BOS = 3, 1, 0
INT_MAX = 0x7FFFFFFF


def _lib():
    from huggingface_asr_b200 import _lib as L

    return L, L.lib()


def _state_after_one_step(B, W, T, V, kind, seed, n_steps=2):
    """A lazy processor advanced n_steps so that r_prev / s_prev / last ids are a realistic mid-decode state."""
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits

    logits, lens, _ = make_encoder_logits(B, T, V, kind, True, seed=seed)
    proc = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
    ids = torch.zeros((B * W, 1), dtype=torch.long, device="cuda")
    beam_scores = torch.zeros(B, W, device="cuda")
    beam_scores[:, 1:] = -1e9
    for n in range(n_steps):
        att = make_attention_scores(B * W, V, n, seed=seed, scale=0.5).cuda()
        out = proc(ids, att)
        cand = (out + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        base = (torch.arange(B, device="cuda") * W).view(B, 1)
        ids = torch.cat([ids[(src + base).view(-1)], tok.view(-1, 1)], dim=1)
        beam_scores = top
    sc = proc.ctc_prefix_scorer
    sel = sc.index_select_state(proc.ctc_states, ids[:, -1].reshape(-1, W))
    return proc, sc, ids, beam_scores.contiguous(), sel


def _dense_and_lists(sc, ids, beam_scores, sel, att, W, w=0.3):
    """(joint, log_psi) of ctcps_score_lazy and (lists, log_psi) of ctcps_score_lazy_topk on the same inputs."""
    L, lib = _lib()
    B, T, V = sc.batch, sc.input_length, sc.odim
    BW = B * W
    ol = ids.shape[1] - 1
    last = ids[:, -1].contiguous()
    r_prev, s_vec = sel[0].contiguous(), sel[1][:, 0].contiguous()
    st = torch.cuda.current_stream().cuda_stream
    ws = sc._workspace(W, 0)
    att_a, att_b = att.clone(), att.clone()
    log_psi = torch.empty((BW, V), device="cuda")
    joint = torch.empty((BW, V), device="cuda")
    x = sc._frame_major()
    L.check(lib.ctcps_score_lazy(x.data_ptr(), sc._ldx, sc._blank_lp.data_ptr(), r_prev.data_ptr(), s_vec.data_ptr(), 1, 0, last.data_ptr(), ol,
                                 B, W, T, V, BLANK, att_a.data_ptr(), 1.0 - w, w, log_psi.data_ptr(), None, joint.data_ptr(), ws.data_ptr(),
                                 ws.numel(), 0, st), "ctcps_score_lazy")
    nl, kk = ctypes.c_int(0), ctypes.c_int(0)
    L.check(lib.ctcps_topk_lists_shape(B, W, V, ctypes.byref(nl), ctypes.byref(kk)), "shape")
    assert kk.value == 2 * W
    lists = torch.full((B, nl.value, kk.value, 2), float("nan"), device="cuda")
    log_psi_t = torch.full((BW, V), float("nan"), device="cuda")
    L.check(lib.ctcps_score_lazy_topk(x.data_ptr(), sc._ldx, r_prev.data_ptr(), s_vec.data_ptr(), last.data_ptr(), ol, B, W, T, V, BLANK,
                                      att_b.data_ptr(), 1.0 - w, w, beam_scores.data_ptr(), log_psi_t.data_ptr(), lists.data_ptr(), ws.data_ptr(),
                                      ws.numel(), 0, st), "ctcps_score_lazy_topk")
    torch.cuda.synchronize()
    assert torch.equal(log_psi_t, log_psi), "log_psi of the fused kernel differs from the dense kernel's"
    return joint, log_psi, lists


def _check_lists(joint, lists, beam_scores, B, W, V):
    """Every tile list = the exact top-K of its (hypothesis group x 512 tokens) block of joint + beam score."""
    K = 2 * W
    key = (joint + beam_scores.view(-1, 1)).view(B, W, V)
    nl = lists.shape[1]
    lk = lists[..., 0]
    li = lists[..., 1].contiguous().view(torch.int32)
    nvt = (V + 511) // 512
    assert nl == nvt and W <= 20, "these checks cover one hypothesis group per tile (W <= 20)"
    for vt in range(nvt):
        v0, v1 = vt * 512, min(V, vt * 512 + 512)
        flat = key[:, :, v0:v1].reshape(B, -1)
        n = flat.shape[1]
        hyp = torch.arange(W, device="cuda").view(-1, 1).expand(W, v1 - v0)
        tok = torch.arange(v0, v1, device="cuda").view(1, -1).expand(W, v1 - v0)
        dense = (hyp * V + tok).reshape(-1)
        # exact order: key descending, dense index ascending (stable sort of index-ordered data)
        m = min(K, n)
        order = torch.sort(flat, dim=1, descending=True, stable=True).indices[:, :m]
        got_key, got_idx = lk[:, vt], li[:, vt]
        assert torch.equal(got_idx[:, :m].long(), dense[order]), f"tile {vt}: candidate indices differ"
        assert torch.equal(got_key[:, :m], torch.gather(flat, 1, order)), f"tile {vt}: keys differ"
        if m < K:
            assert (got_idx[:, m:] == INT_MAX).all() and torch.isinf(got_key[:, m:]).all()


@pytest.mark.parametrize("B,W,T,V,kind", [
    (3, 10, 100, 1200, "peaky"),   # 3 tiles, the last one partly filled
    (2, 7, 61, 516, "flat"),       # odd beam (group padded to 8), second tile holds 4 tokens: fewer valid threads than K -> fallback
    (2, 20, 90, 260, "peaky"),     # W = 20: one group of 20 hypotheses, K = 40, half-empty tile
    (4, 10, 373, 5000, "peaky"),   # BASELINE vocabulary and length
    (2, 1, 50, 300, "flat"),       # greedy-width beam: K = 2
    (3, 4, 40, 64, "peaky"),       # tiny vocabulary: 16 valid threads
])
def test_tile_lists_are_the_exact_top_2w(B, W, T, V, kind):
    from huggingface_asr_b200.synthetic import make_attention_scores

    proc, sc, ids, beam_scores, sel = _state_after_one_step(B, W, T, V, kind, seed=77 + W)
    att = make_attention_scores(B * W, V, 5, seed=3, scale=0.5).cuda()
    joint, log_psi, lists = _dense_and_lists(sc, ids, beam_scores, sel, att, W)
    _check_lists(joint, lists, beam_scores, B, W, V)


def test_tile_lists_with_ties_and_minus_infinity():
    """Constant decoder rows (every key of a hypothesis ties -> more than the candidate cap reach the threshold: the argmax
    fallback) and -inf decoder scores (suppressed tokens): still the exact (key, index) order."""
    B, W, T, V = 2, 4, 60, 1024
    proc, sc, ids, beam_scores, sel = _state_after_one_step(B, W, T, V, "flat", seed=5)
    att = torch.zeros(B * W, V, device="cuda")
    att[1::2, ::3] = float("-inf")
    att[2, 100:140] = 1.0
    joint, log_psi, lists = _dense_and_lists(sc, ids, beam_scores, sel, att, W)
    _check_lists(joint, lists, beam_scores, B, W, V)
    # a state where every CTC score is identical as well: step 0 of a flat-prior utterance is not, so force constant keys
    att2 = torch.full((B * W, V), -3.0, device="cuda")
    joint, log_psi, lists = _dense_and_lists(sc, ids, beam_scores, sel, att2, W, w=0.0)
    assert (joint[:, 5] == joint[:, 900]).all()
    _check_lists(joint, lists, beam_scores, B, W, V)


def _beam_buffers(B, W, V, L, maxlen, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ids_cur = torch.randint(5, V, (B * W, maxlen), device="cuda", generator=g)
    pool_scores = torch.full((B, W), float("-inf"), device="cuda")
    pool_scores[0, : W // 2] = -5.0 - torch.arange(W // 2, device="cuda")  # a half-filled pool
    pool_lens = torch.zeros(B, W, dtype=torch.long, device="cuda")
    pool_seqs = torch.zeros(B, W, maxlen, dtype=torch.long, device="cuda")
    done = torch.zeros(B, dtype=torch.uint8, device="cuda")
    return ids_cur, pool_scores, pool_lens, pool_seqs, done


@pytest.mark.parametrize("B,W,T,V", [(3, 10, 100, 1200), (2, 20, 90, 260), (5, 3, 40, 64)])
def test_beam_step_over_lists_equals_the_dense_beam_step(B, W, T, V):
    """Same outputs as ctcps_beam_step on the dense joint scores: beam scores, rows, pool, done flags, selection ids."""
    from huggingface_asr_b200.synthetic import make_attention_scores

    L, lib = _lib()
    proc, sc, ids, beam_scores, sel = _state_after_one_step(B, W, T, V, "peaky", seed=11 + W)
    att = make_attention_scores(B * W, V, 9, seed=4, scale=0.5).cuda()
    att[:, EOS] += 6.0  # make eos competitive so that hypotheses are finalised into the pool
    att = torch.log_softmax(att, -1)
    joint, log_psi, lists = _dense_and_lists(sc, ids, beam_scores, sel, att, W)
    maxlen, Lcur = 16, ids.shape[1]
    n = ctypes.c_size_t(0)
    lib.ctcps_beam_step_workspace_bytes(B, W, ctypes.byref(n))
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for fused in (False, True):
        ids_cur, pool_scores, pool_lens, pool_seqs, done = _beam_buffers(B, W, V, Lcur, maxlen, seed=2)
        ids_cur[:, :Lcur] = ids
        ids_next = torch.zeros_like(ids_cur)
        bs = beam_scores.clone()
        ws = torch.zeros((n.value + 15) // 16 * 2, dtype=torch.int64, device="cuda")
        best = torch.zeros((B, W), dtype=torch.long, device="cuda")
        last = torch.zeros((B * W,), dtype=torch.long, device="cuda")
        common = (bs.data_ptr(), ids_cur.data_ptr(), ids_next.data_ptr(), maxlen, Lcur, B, W, V, EOS, BLANK, float(Lcur), pool_scores.data_ptr(),
                  pool_lens.data_ptr(), pool_seqs.data_ptr(), maxlen, done.data_ptr(), ws.data_ptr(), ws.numel() * 8, None, 0, 0)
        if fused:
            L.check(lib.ctcps_beam_step_lists(lists.data_ptr(), lists.shape[1], *common, best.data_ptr(), last.data_ptr(), st), "ctcps_beam_step_lists")
        else:
            L.check(lib.ctcps_beam_step(joint.data_ptr(), *common, best.data_ptr(), st), "ctcps_beam_step")
        torch.cuda.synchronize()
        outs.append((bs, ids_next[:, : Lcur + 1].clone(), pool_scores, pool_lens, pool_seqs, done, best, last))
    a, b = outs
    for k, name in enumerate(["beam_scores", "ids_next", "pool_scores", "pool_lens", "pool_seqs", "done", "best_ids"]):
        assert torch.equal(a[k], b[k]), f"{name} differs between the dense and the list beam step"
    assert (b[3] > 0).any(), "the case should finalise at least one hypothesis"
    assert torch.equal(b[7], (b[6] % V).view(-1))


@pytest.mark.parametrize("use_beam_idx", [False, True])
@pytest.mark.parametrize("B,W,T,V", [(8, 10, 248, 5000), (3, 20, 120, 1000), (4, 5, 30, 260)])
def test_native_loop_with_fused_topk_equals_the_unfused_loop(B, W, T, V, use_beam_idx):
    """Whole decodes: hypotheses, lengths and scores identical bit for bit, same number of steps.  The (4, 5, 30, 260) decode is cut
    at max_length > T: prefixes outgrow the utterance and the loop must fall back to the dense step for those steps."""
    from huggingface_asr_b200.beam_search import joint_beam_search_native
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import SyntheticDecoder, make_attention_scores, make_encoder_logits

    logits, lens, tr = make_encoder_logits(B, T, V, "peaky", True, seed=23)
    max_length = 48
    if T < max_length:
        dec = lambda ids, n: make_attention_scores(B * W, V, n, seed=3, scale=0.5).cuda()  # noqa: E731  never emits eos early
    else:
        dec = SyntheticDecoder(tr, W, V, max_length, seed=1, device="cuda")
    outs = []
    for fuse in (False, True):
        proc = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False,
                                          use_beam_idx=use_beam_idx)
        outs.append(joint_beam_search_native(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=max_length, device=torch.device("cuda"),
                                             done_check_lag=0, fuse_topk=fuse))
    a, b = outs
    assert a.steps == b.steps
    assert torch.equal(a.sequences, b.sequences) and torch.equal(a.lengths, b.lengths)
    assert torch.equal(a.scores, b.scores), f"scores differ by {(a.scores - b.scores).abs().max().item()}"


@pytest.mark.parametrize("B,W,T,V,lag", [(12, 10, 96, 1000, 1), (6, 20, 120, 516, 0), (9, 4, 60, 2048, 2)])
def test_skipping_finished_utterances_changes_nothing_that_is_returned(B, W, T, V, lag):
    """The native loop does not score utterances whose beam search has finished (ctcps_set_skip_done, default on).  Ragged
    lengths make the utterances finish at different steps; n-best sequences, lengths, scores and the number of steps must be
    identical bit for bit with the switch off, for every done-check lag."""
    from huggingface_asr_b200 import _lib
    from huggingface_asr_b200.beam_search import joint_beam_search_native
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import SyntheticDecoder, make_encoder_logits

    logits, lens, tr = make_encoder_logits(B, T, V, "peaky", True, seed=31)
    assert len({len(t) for t in tr}) > 1, "the transcripts must differ in length"
    L = _lib.lib()
    prev = L.ctcps_set_skip_done(-1)
    outs = []
    try:
        for on in (0, 1):
            L.ctcps_set_skip_done(on)
            dec = SyntheticDecoder(tr, W, V, 64, seed=2, device="cuda")
            proc = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
            outs.append(joint_beam_search_native(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=64, device=torch.device("cuda"),
                                                 done_check_lag=lag, fuse_topk=True, num_return_sequences=min(W, 3)))
    finally:
        L.ctcps_set_skip_done(prev)
    a, b = outs
    assert a.steps == b.steps
    assert torch.equal(a.sequences, b.sequences) and torch.equal(a.lengths, b.lengths)
    assert torch.equal(a.scores, b.scores), f"scores differ by {(a.scores - b.scores).abs().max().item()}"
    assert torch.equal(a.nbest_sequences, b.nbest_sequences) and torch.equal(a.nbest_lengths, b.nbest_lengths)
    assert torch.equal(a.nbest_scores, b.nbest_scores)
    assert (a.lengths.cpu() == torch.tensor([len(t) - 1 for t in tr])).all()


@pytest.mark.parametrize("B,W,T,V,kind,n_steps", [
    (6, 10, 100, 1200, "peaky", 2),   # ragged lengths 60 .. 100 frames: up to 5 of 13 chunks are padding
    (5, 4, 61, 516, "flat", 1),
    (4, 20, 90, 260, "peaky", 3),
    (3, 3, 40, 64, "peaky", 0),       # first step (ol = 0: the x[0] term comes from chunk 0)
])
def test_leaving_out_padded_frames_is_bit_identical(B, W, T, V, kind, n_steps):
    """ctcps_score_lazy_lens / _topk_active with the utterance lengths do not stream the chunks past an utterance's last frame:
    exp(logzero) is exactly 0, so log_psi, joint scores and the tile lists must equal those of the calls without lengths bit for
    bit -- including an utterance of length 0 and one that ends exactly on a chunk boundary."""
    L, lib = _lib()
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits

    logits, lens, _ = make_encoder_logits(B, T, V, kind, True, seed=71)
    lens[0] = 0
    lens[1] = 48 if T > 48 else 8     # a multiple of the 8-frame chunk
    proc = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
    sc = proc.ctc_prefix_scorer
    ids = torch.zeros((B * W, 1), dtype=torch.long, device="cuda")
    beam_scores = torch.zeros(B, W, device="cuda")
    beam_scores[:, 1:] = -1e9
    for n in range(n_steps):
        att = make_attention_scores(B * W, V, n, seed=5, scale=0.5).cuda()
        out = proc(ids, att)
        cand = (out + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        base = (torch.arange(B, device="cuda") * W).view(B, 1)
        ids = torch.cat([ids[(idx // V + base).view(-1)], (idx % V).view(-1, 1)], dim=1)
        beam_scores = top
    beam_scores = beam_scores.contiguous()
    if n_steps > 0:
        sel = sc.index_select_state(proc.ctc_states, ids[:, -1].reshape(-1, W))
        r_prev, s_ptr = sel[0].contiguous(), sel[1][:, 0].contiguous()
    else:
        r_prev, s_ptr = sc.initial_state(W), None
    att = make_attention_scores(B * W, V, n_steps, seed=5, scale=0.5).cuda()
    BW, ol = B * W, ids.shape[1] - 1
    last = ids[:, -1].contiguous()
    st = torch.cuda.current_stream().cuda_stream
    ws = sc._workspace(W, 0)
    x = sc._frame_major()
    xl = sc._score_lens()
    assert xl is not None
    nl, kk = ctypes.c_int(0), ctypes.c_int(0)
    L.check(lib.ctcps_topk_lists_shape(B, W, V, ctypes.byref(nl), ctypes.byref(kk)), "shape")
    outs = []
    for lens_ptr in (None, xl.data_ptr()):
        a1, a2 = att.clone(), att.clone()
        log_psi, ts, joint = (torch.full((BW, V), float("nan"), device="cuda") for _ in range(3))
        L.check(lib.ctcps_score_lazy_lens(x.data_ptr(), sc._ldx, sc._blank_lp.data_ptr(), lens_ptr, r_prev.data_ptr(),
                                          None if s_ptr is None else s_ptr.data_ptr(), 1, 0, last.data_ptr(), ol, B, W, T, V, BLANK,
                                          a1.data_ptr(), 0.7, 0.3, log_psi.data_ptr(), ts.data_ptr(), joint.data_ptr(), ws.data_ptr(),
                                          ws.numel(), 0, st), "ctcps_score_lazy_lens")
        res = [log_psi, ts, joint]
        if V % 4 == 0:
            lists = torch.full((B, nl.value, kk.value, 2), float("nan"), device="cuda")
            lp2 = torch.full((BW, V), float("nan"), device="cuda")
            L.check(lib.ctcps_score_lazy_topk_active(x.data_ptr(), sc._ldx, r_prev.data_ptr(), None if s_ptr is None else s_ptr.data_ptr(),
                                                     last.data_ptr(), ol, B, W, T, V, BLANK, a2.data_ptr(), 0.7, 0.3, beam_scores.data_ptr(),
                                                     None, lens_ptr, lp2.data_ptr(), lists.data_ptr(), ws.data_ptr(), ws.numel(), 0, st),
                    "ctcps_score_lazy_topk_active")
            res += [lists, lp2]
        torch.cuda.synchronize()
        outs.append(res)
    for a, b in zip(*outs):
        assert not torch.isnan(a).any()
        assert torch.equal(a, b)


@pytest.mark.parametrize("B,W,T,V,kind", [(6, 10, 200, 1000, "peaky"), (4, 20, 300, 516, "peaky"), (5, 4, 120, 2048, "flat"), (3, 3, 40, 64, "peaky")])
def test_streaming_only_the_frames_with_nonzero_weight_is_bit_identical(B, W, T, V, kind):
    """ctcps_set_frame_window (default on): the lazy scoring kernel streams only the chunks in which exp(r_sum - offset) is nonzero
    for some hypothesis of the group.  Whole decodes (native loop: n-best, scores) and the processor's dense outputs at every
    step must be bit-identical with the switch off -- and the switch must actually save chunks on a long utterance."""
    from huggingface_asr_b200 import _lib
    from huggingface_asr_b200.beam_search import joint_beam_search_native
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import SyntheticDecoder, make_attention_scores, make_encoder_logits

    logits, lens, tr = make_encoder_logits(B, T, V, kind, True, seed=47)
    L = _lib.lib()
    prev = L.ctcps_set_frame_window(-1)
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    outs, dense, chunks = [], [], []
    try:
        for on in (0, 1):
            L.ctcps_set_frame_window(on)
            L.ctcps_set_stream_counter(counter.data_ptr())
            counter.zero_()
            if kind == "peaky":
                dec = SyntheticDecoder(tr, W, V, 64, seed=2, device="cuda")
            else:
                dec = lambda ids, n: make_attention_scores(B * W, V, n, seed=3, scale=0.5).cuda()  # noqa: E731
            proc = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
            outs.append(joint_beam_search_native(proc, dec, B, W, V, BOS, EOS, BLANK, max_length=min(64, T // 2), device=torch.device("cuda"),
                                                 done_check_lag=0, fuse_topk=True, num_return_sequences=min(W, 3)))
            torch.cuda.synchronize()
            chunks.append(int(counter.item()))
            # the processor's own __call__ (dense joint scores), a few steps of a plain beam update
            proc = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
            ids = torch.zeros((B * W, 1), dtype=torch.long, device="cuda")
            beam_scores = torch.zeros(B, W, device="cuda")
            beam_scores[:, 1:] = -1e9
            steps = []
            for n in range(6):
                att = make_attention_scores(B * W, V, n, seed=5, scale=0.5).cuda()
                out = proc(ids, att)
                steps.append(out.clone())
                top, idx = (out + beam_scores.view(-1, 1)).view(B, W * V).topk(W, dim=1)
                base = (torch.arange(B, device="cuda") * W).view(B, 1)
                ids = torch.cat([ids[(idx // V + base).view(-1)], (idx % V).view(-1, 1)], dim=1)
                beam_scores = top
            dense.append(steps)
    finally:
        L.ctcps_set_stream_counter(None)
        L.ctcps_set_frame_window(prev)
    a, b = outs
    assert a.steps == b.steps
    assert torch.equal(a.sequences, b.sequences) and torch.equal(a.lengths, b.lengths) and torch.equal(a.scores, b.scores)
    assert torch.equal(a.nbest_sequences, b.nbest_sequences) and torch.equal(a.nbest_scores, b.nbest_scores)
    for n, (x, y) in enumerate(zip(*dense)):
        assert torch.equal(x, y), f"processor output of step {n} differs by {(x - y).abs().max().item()}"
    print(f"chunks streamed by the native decode: all frames {chunks[0]}, windowed {chunks[1]}")
    assert chunks[1] <= chunks[0]
    if T >= 200:
        assert chunks[1] < 0.8 * chunks[0], "the window should leave out a fair share of a long utterance"
