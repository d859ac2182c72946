"""GPU parity of the pre-beam (candidate) decode step -- SURVEY.md 8(f) N2 -- through the C ABI:
the reference SCORER under ESPnet's pre-beam policy (golden traces), the CPU oracle on larger seeded cases, and the
internal consistency of the three forms of the step (materialised partial state, lazy candidates with dense outputs,
sparse candidates + ctcps_beam_step_candidates).  Run on the B200 box with -m gpu."""
import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu

BLANK, EOS, BOS = 3, 1, 0


def _proc(logits, lens, w, W, S, use_beam_idx, materialize=False, **kw):
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor

    return CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, w, W, -1, False, 1.0, materialize_state=materialize,
                                      pre_beam_size=S, use_beam_idx=use_beam_idx, **kw)


@pytest.mark.parametrize("name", parity.PREBEAM_CASES)
@pytest.mark.parametrize("materialize", [False, True])
def test_prebeam_vs_reference_golden(name, materialize):
    worst = parity.replay_prebeam(lambda lg, ln, w, W, S, ubi: _proc(lg, ln, w, W, S, ubi, materialize), "cuda", name,
                                  teacher_forced=name.endswith("tokens_only"))
    print(name, "materialized" if materialize else "lazy", worst)


@pytest.mark.parametrize("name", parity.PREBEAM_CASES)
def test_sparse_fused_decode_vs_reference_golden(name):
    """score_candidates + ctcps_beam_step_candidates (no (BW,V) tensor anywhere) gives the golden 1-best."""
    from huggingface_asr_b200.beam_search import joint_beam_search_fused
    from huggingface_asr_b200.synthetic import make_attention_scores

    g = parity.load(name)
    W, S, seed = int(g["W"]), int(g["S"]), int(g["seed"])
    logits, lens = torch.from_numpy(g["logits"]).cuda(), torch.from_numpy(g["lens"]).cuda()
    B, T, V = logits.shape
    if not bool(g["use_beam_idx"]):
        pytest.skip("token-only selection with candidates produces exact ties at 3e9 (see parity.replay_prebeam)")
    proc = _proc(logits, lens, float(g["ctc_weight"]), W, S, bool(g["use_beam_idx"]))
    out = joint_beam_search_fused(proc, lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5).cuda(), B, W, V, BOS,
                                  EOS, BLANK, max_length=int(g["max_length"]), device="cuda")
    assert out.steps == int(g["steps"])
    assert (out.sequences.cpu().numpy() == g["seq"]).all(), f"{name}: 1-best differs"
    assert (out.lengths.cpu().numpy() == g["len"]).all()
    assert np.abs(out.scores.cpu().numpy() - g["score"]).max() <= 1e-4


def _seeded(B, W, T, V, kind="peaky", ragged=True, seed=71):
    from huggingface_asr_b200.synthetic import make_encoder_logits

    return make_encoder_logits(B, T, V, kind, ragged, seed=seed)


@pytest.mark.parametrize("B,W,T,V,S", [(3, 10, 96, 1000, 15), (2, 20, 120, 777, 40), (4, 7, 64, 5000, 32)])
def test_prebeam_decode_vs_oracle(B, W, T, V, S):
    """Whole decodes on shapes with several V tiles / W = 20 / S > 32: per-step joint scores against the oracle under the
    torch harness, then the sparse fused harness must give the same 1-best."""
    from huggingface_asr_b200.beam_search import joint_beam_search, joint_beam_search_fused
    from huggingface_asr_b200.synthetic import make_attention_scores
    from oracle import oracle as orc

    logits, lens, _ = _seeded(B, W, T, V)
    cpu = orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W, pre_beam_size=S)
    trace = []

    class Rec:
        use_beam_idx = True

        def set_beam_idx(self, bi):
            cpu.set_beam_idx(bi)

        def __call__(self, ids, scores):
            out = cpu(ids, scores)
            trace.append((ids.clone(), out.clone()))
            return out

    def att(dev):
        return lambda ids, n: make_attention_scores(B * W, V, n, seed=9, scale=0.5).to(dev)

    oc = joint_beam_search(Rec(), att("cpu"), B, W, V, BOS, EOS, BLANK, max_length=20)

    gpu = _proc(logits.cuda(), lens.cuda(), 0.3, W, S, True)
    step = [0]

    class Chk:
        use_beam_idx = True

        def set_beam_idx(self, bi):
            gpu.set_beam_idx(bi)

        def __call__(self, ids, scores):
            ref_ids, ref_out = trace[step[0]]
            assert (ids.cpu() == ref_ids).all(), f"step {step[0]}: decode diverged"
            out = gpu(ids, scores)
            parity.assert_parity(out, ref_out, f"step {step[0]} joint")
            step[0] += 1
            return out

    og = joint_beam_search(Chk(), att("cuda"), B, W, V, BOS, EOS, BLANK, max_length=20, device="cuda")
    assert og.steps == oc.steps and (og.sequences.cpu() == oc.sequences).all()
    of = joint_beam_search_fused(_proc(logits.cuda(), lens.cuda(), 0.3, W, S, True), att("cuda"), B, W, V, BOS, EOS, BLANK,
                                 max_length=20, device="cuda")
    assert of.steps == oc.steps and (of.sequences.cpu() == oc.sequences).all(), "sparse fused decode differs from the oracle"
    assert (of.scores.cpu() - oc.scores).abs().max() <= 1e-4


def test_candidate_scores_equal_those_of_the_full_vocabulary_step():
    """On the same state the joint scores of the candidates are those of the full-vocabulary lazy step (same lin stream, same
    posteriors; only the order of the sum over frames differs: warp-wide tree here, frame order there): <= 2e-5."""
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores

    B, W, T, V, S = 3, 10, 96, 1000, 24
    logits, lens, _ = _seeded(B, W, T, V, seed=72)
    full = CTCRescorerLogitsProcessor(logits.cuda(), lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
    cand = _proc(logits.cuda(), lens.cuda(), 0.3, W, S, False)  # token-only selection in both: identical states
    ids = torch.full((B * W, 1), BOS, dtype=torch.long, device="cuda")
    for n in range(3):
        att = make_attention_scores(B * W, V, n, seed=4, scale=0.5).cuda()
        jf = full(ids, att.clone())
        cid, cj = cand.score_candidates(ids, att.clone())
        assert cid.shape == (B * W, S) and (cid.sort(1).values[:, 1:] != cid.sort(1).values[:, :-1]).all(), "ids not unique"
        ref = jf.gather(1, cid)
        fin = ref > -1e9
        assert bool((cj[~fin] <= -1e9).all()) and float((cj[fin] - ref[fin]).abs().max()) <= 2e-5, f"step {n}"
        # both continue with tokens that hypothesis 0 of the utterance scored (the token-only selection reads its column)
        nxt = cid.view(B, W, S)[:, 0, :W].reshape(-1, 1)
        ids = torch.cat([ids, nxt], 1)


def test_prebeam_topk_matches_torch_sort():
    from huggingface_asr_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(3)
    # register fast path (V % 4 == 0, V <= 8192), general radix path (odd V, V > 8192), S up to 64
    for BW, V, S in [(37, 5000, 32), (5, 129, 40), (64, 1000, 1), (3, 70, 64), (9, 5000, 64), (4, 9000, 15), (6, 8192, 20), (5, 64, 15)]:
        att = torch.randn(BW, V, generator=g).cuda()
        att[0, 5:9] = att[0, 4]          # ties: lower id first
        att[1 % BW, 10] = float("-inf")  # masked token
        if BW > 2:
            att[2] = -7.25               # a constant row: > TOPK_CAP elements at the threshold -> radix select, id order
        if BW > 3:
            att[3] = torch.log_softmax(torch.randn(V, generator=g) * 0.5, -1).round(decimals=1).cuda()  # heavy ties at the boundary
        if BW > 4:
            att[4, ::2] = 3.5            # half the row shares the best score
        ref = att.clone()
        ref[:, BLANK] = -1e10
        ids = torch.empty((BW, S), dtype=torch.long, device="cuda")
        val = torch.empty((BW, S), dtype=torch.float32, device="cuda")
        _lib.check(L.ctcps_prebeam_topk(att.data_ptr(), BW, V, BLANK, S, ids.data_ptr(), val.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "topk")
        order = torch.sort(ref, dim=1, descending=True, stable=True)
        assert torch.equal(ids, order.indices[:, :S]), (BW, V, S)
        assert torch.equal(val, order.values[:, :S])
        assert torch.equal(att, ref), "scores[:, pad] = logzero must happen in place"


def test_token_major_layout_round_trips():
    from huggingface_asr_b200.decoding.ctc_scorer import CTCPrefixScoreTH

    B, W, T, V = 2, 3, 37, 131
    logits, lens, _ = _seeded(B, W, T, V, seed=73)
    a = CTCPrefixScoreTH.from_logits(logits.cuda(), lens.cuda(), BLANK, EOS)
    b = CTCPrefixScoreTH.from_logits(logits.cuda(), lens.cuda(), BLANK, EOS, token_major=True)
    assert b._x is None and b._xt.shape == (B, V, b._ldt) and b._ldt % 4 == 0
    assert torch.equal(b._xt[:, :, :T].transpose(1, 2), a._x[:, :, :V])
    assert bool((b._xt[:, :, T:] == 0).all())
    assert torch.equal(a.x, b.x)  # rebuilt frame-major copy
    assert torch.equal(a._token_major(), b._xt)


def test_prebeam_argument_errors():
    logits, lens, _ = _seeded(2, 3, 20, 40, seed=74)
    with pytest.raises(ValueError):
        _proc(logits.cuda(), lens.cuda(), 0.3, 3, 65, True)
    p = _proc(logits.cuda(), lens.cuda(), 0.3, 3, 0, False)
    with pytest.raises(RuntimeError):
        p.score_candidates(torch.zeros((6, 1), dtype=torch.long, device="cuda"), torch.zeros((6, 40), device="cuda"))


# ------------------------------------------------------------------------------------------------
# native decode-step driver (ctcps_decode_step): one host call per step, same kernels, same results
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in parity.PREBEAM_CASES if not n.endswith("tokens_only")])
@pytest.mark.parametrize("lag", [0, 1])
def test_native_loop_prebeam_vs_reference_golden(name, lag):
    from huggingface_asr_b200.beam_search import joint_beam_search_native
    from huggingface_asr_b200.synthetic import make_attention_scores

    g = parity.load(name)
    W, S, seed = int(g["W"]), int(g["S"]), int(g["seed"])
    logits, lens = torch.from_numpy(g["logits"]).cuda(), torch.from_numpy(g["lens"]).cuda()
    B, T, V = logits.shape
    proc = _proc(logits, lens, float(g["ctc_weight"]), W, S, True)
    out = joint_beam_search_native(proc, lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5).cuda(), B, W, V, BOS,
                                   EOS, BLANK, max_length=int(g["max_length"]), device="cuda", done_check_lag=lag)
    assert out.steps >= int(g["steps"]) and out.steps <= int(g["steps"]) + lag  # lag > 0 may run extra (frozen) steps
    assert (out.sequences.cpu().numpy() == g["seq"]).all(), f"{name}: 1-best differs"
    assert (out.lengths.cpu().numpy() == g["len"]).all()
    assert np.abs(out.scores.cpu().numpy() - g["score"]).max() <= 1e-4


def test_native_loop_full_vocabulary_vs_reference_golden():
    """S = 0: the lazy full-vocabulary step (the bench line) through ctcps_decode_step gives the reference's 1-best."""
    from huggingface_asr_b200.beam_search import joint_beam_search_native
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import make_attention_scores

    g = parity.load("decode_1best")
    for i in range(3):
        logits, lens = torch.from_numpy(g[f"d{i}_logits"]).cuda(), torch.from_numpy(g[f"d{i}_lens"]).cuda()
        W, seed = int(g[f"d{i}_W"]), int(g[f"d{i}_seed"])
        B, T, V = logits.shape
        proc = CTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, 0.3, W, -1, False, 1.0, materialize_state=False)
        timing = []
        out = joint_beam_search_native(proc, lambda ids, n: make_attention_scores(B * W, V, n, seed=seed, scale=0.5).cuda(), B, W, V,
                                       BOS, EOS, BLANK, max_length=int(g[f"d{i}_max_length"]), device="cuda", done_check_lag=0,
                                       score_timing=timing)
        assert out.steps == int(g[f"d{i}_steps"])
        assert (out.sequences.cpu().numpy() == g[f"d{i}_seq"]).all(), f"decode {i}: 1-best differs from the reference"
        assert np.abs(out.scores.cpu().numpy() - g[f"d{i}_score"]).max() <= 1e-4
        from huggingface_asr_b200.beam_search import resolve_score_timing

        torch.cuda.synchronize()
        ms = resolve_score_timing(timing)
        assert len(ms) == out.steps and all(m > 0 for m in ms)


@pytest.mark.parametrize("S", [0, 15, 40])
def test_native_loop_equals_fused_loop(S):
    """Same kernels behind one host call per step: identical hypotheses, scores bit for bit, on a C1-like shape."""
    from huggingface_asr_b200.beam_search import joint_beam_search_fused, joint_beam_search_native
    from huggingface_asr_b200.synthetic import SyntheticDecoder, make_encoder_logits

    B, W, T, V = 8, 10, 120, 5000
    logits, lens, transcripts = make_encoder_logits(B, T, V, "peaky", True, seed=81)
    dec = SyntheticDecoder(transcripts, W, V, 40, seed=3, device="cuda")
    a = joint_beam_search_fused(_proc(logits.cuda(), lens.cuda(), 0.3, W, S, None), dec, B, W, V, BOS, EOS, BLANK, max_length=40,
                                device="cuda")
    b = joint_beam_search_native(_proc(logits.cuda(), lens.cuda(), 0.3, W, S, None), dec, B, W, V, BOS, EOS, BLANK, max_length=40,
                                 device="cuda", done_check_lag=0)
    assert a.steps == b.steps and torch.equal(a.sequences, b.sequences) and torch.equal(a.scores, b.scores)
    want = torch.tensor([len(t) - 1 for t in transcripts])  # hypotheses are stored without their eos
    assert (a.lengths.cpu() == want).all(), "the decode does not recover the planted transcripts"


def test_prebeam_edge_shapes_vs_oracle():
    """Shapes that take the other code paths: V % 4 != 0 (radix top-k, scalar loads), S = 64 (two candidate passes per warp),
    W = 20 (two hypothesis groups), T > 384 (several row segments per candidate), ragged lengths with a very short utterance."""
    from huggingface_asr_b200.beam_search import joint_beam_search, joint_beam_search_native
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits
    from oracle import oracle as orc

    B, W, T, V, S = 2, 20, 520, 203, 64
    logits, lens, _ = make_encoder_logits(B, T, V, "peaky", True, seed=91)
    lens[1] = 40  # much shorter than T, but longer than the decode: past its end every score ties at -3e9 and CPU / CUDA topk
    #               break exact ties differently (that regime is test_prebeam_prefix_longer_than_the_utterance's)
    cpu = orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W, pre_beam_size=S)
    trace = []

    class Rec:
        use_beam_idx = True

        def set_beam_idx(self, bi):
            cpu.set_beam_idx(bi)

        def __call__(self, ids, scores):
            out = cpu(ids, scores)
            trace.append((ids.clone(), out.clone()))
            return out

    def att(dev):
        return lambda ids, n: make_attention_scores(B * W, V, n, seed=19, scale=0.5).to(dev)

    oc = joint_beam_search(Rec(), att("cpu"), B, W, V, BOS, EOS, BLANK, max_length=12)
    gpu = _proc(logits.cuda(), lens.cuda(), 0.3, W, S, True)
    step = [0]

    class Chk:
        use_beam_idx = True

        def set_beam_idx(self, bi):
            gpu.set_beam_idx(bi)

        def __call__(self, ids, scores):
            ref_ids, ref_out = trace[step[0]]
            assert (ids.cpu() == ref_ids).all(), f"step {step[0]}: decode diverged"
            out = gpu(ids, scores)
            parity.assert_parity(out, ref_out, f"step {step[0]} joint")
            step[0] += 1
            return out

    og = joint_beam_search(Chk(), att("cuda"), B, W, V, BOS, EOS, BLANK, max_length=12, device="cuda")
    assert og.steps == oc.steps and (og.sequences.cpu() == oc.sequences).all()
    on = joint_beam_search_native(_proc(logits.cuda(), lens.cuda(), 0.3, W, S, True), att("cuda"), B, W, V, BOS, EOS, BLANK,
                                  max_length=12, device="cuda", done_check_lag=0)
    assert on.steps == oc.steps and (on.sequences.cpu() == oc.sequences).all()
    assert (on.scores.cpu() - oc.scores).abs().max() <= 1e-4


def test_prebeam_prefix_longer_than_the_utterance():
    """output_length > T: the reference returns logzero everywhere (:138-145); the candidate path must do the same, and the
    steps around it (start == T - 1, start == T) must match the oracle."""
    from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits
    from oracle import oracle as orc

    B, W, T, V, S = 1, 3, 6, 24, 5
    logits, lens, _ = make_encoder_logits(B, T, V, "flat", False, seed=92)
    cpu = orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W, pre_beam_size=S)
    gpu = _proc(logits.cuda(), lens.cuda(), 0.3, W, S, True)
    ids = torch.full((B * W, 1), BOS, dtype=torch.long)
    g = torch.Generator().manual_seed(1)
    for n in range(T + 3):
        a = make_attention_scores(B * W, V, n, seed=23, scale=0.5)
        oc = cpu(ids, a.clone())
        og = gpu(ids.cuda(), a.cuda())
        parity.assert_parity(og, oc, f"prefix length {n}")
        parity.assert_parity(gpu.ctc_states[1], cpu.ctc_states[1], f"log_psi at prefix length {n}")
        # every beam keeps its own best-scored candidate (never eos / blank): beam_idx = identity
        tok = torch.sort(a, dim=1, descending=True, stable=True).indices
        nxt = torch.stack([next(t for t in row.tolist() if t not in (EOS, BLANK)) * torch.ones((), dtype=torch.long) for row in tok])
        bi = torch.arange(B * W)
        cpu.set_beam_idx(bi)
        gpu.set_beam_idx(bi.cuda())
        ids = torch.cat([ids, nxt.view(-1, 1)], 1)
