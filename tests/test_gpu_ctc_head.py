"""GPU: N4 -- the CTC head at fp32 accuracy on the tensor cores: the hand-written tcgen05 kernel (csrc/ctcps_head.cu: 3xFP16
UMMA on power-of-two-scaled operands, bias + softmax statistics in the drain, TMA stores, streaming normalisation) and round 1's
library form (TF32 operand split + one stacked-K cuBLAS GEMM + K-a), and the processor built from encoder hidden states.  Reference: Wav2Vec2ForCTC.lm_head
(src/reguler/e_branchformer.py:245-252) -> F.log_softmax (src/decoding/ctc_scorer.py:279) -> padding (:39-46).
Run on the B200 box with -m gpu."""
import pytest
import torch

pytestmark = pytest.mark.gpu

BLANK, EOS, BOS = 3, 1, 0


def test_split_is_exact():
    from huggingface_asr_b200.ctc_head import split_tf32

    g = torch.Generator().manual_seed(1)
    x = (torch.randn(37, 64, generator=g) * torch.logspace(-6, 6, 64)).cuda()
    a, w = split_tf32(x, False), split_tf32(x, True)
    hi, lo = a[:, :64], a[:, 64:128]
    assert torch.equal(a[:, 128:], hi) and torch.equal(w[:, :64], lo) and torch.equal(w[:, 64:128], hi) and torch.equal(w[:, 128:], hi)
    assert bool(((hi.view(torch.int32) & 0x1FFF) == 0).all()), "hi must be exactly representable in TF32"
    assert torch.equal(hi + lo, x), "hi + lo must reproduce x exactly"
    assert float((lo.abs() / x.abs()).max()) <= 2.0 ** -11


def test_split_hi_lo_is_exact():
    from huggingface_asr_b200.ctc_head import split_hi_lo

    g = torch.Generator().manual_seed(1)
    x = (torch.randn(37, 64, generator=g) * torch.logspace(-6, 6, 64)).cuda()
    hi, lo = split_hi_lo(x)
    assert bool(((hi.view(torch.int32) & 0x1FFF) == 0).all()), "hi must be exactly representable in TF32"
    assert torch.equal(hi + lo, x), "hi + lo must reproduce x exactly"
    assert float((lo.abs() / x.abs()).max()) <= 2.0 ** -11


@pytest.mark.parametrize("impl", ["tcgen05", "cublas"])
def test_head_has_fp32_accuracy_where_single_pass_tf32_does_not(impl):
    from huggingface_asr_b200.ctc_head import CTCHead
    from huggingface_asr_b200.synthetic import make_encoder_hidden

    hidden, weight, bias, _, _ = make_encoder_hidden(8, 120, 5000, 512, seed=5)
    ref = (hidden.double().reshape(-1, 512) @ weight.double().t() + bias.double()).view(8, 120, 5000)
    head = CTCHead(weight.cuda(), bias.cuda(), implementation=impl)
    out = head(hidden.cuda()).cpu().double()
    err3 = float((out - ref).abs().max())
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    err_fp32 = float((torch.addmm(bias.cuda(), hidden.cuda().view(-1, 512), weight.cuda().t()).cpu().double().view_as(ref) - ref).abs().max())
    torch.backends.cuda.matmul.allow_tf32 = True
    err1 = float((torch.addmm(bias.cuda(), hidden.cuda().view(-1, 512), weight.cuda().t()).cpu().double().view_as(ref) - ref).abs().max())
    torch.backends.cuda.matmul.allow_tf32 = prev
    print(f"{impl}: max |logit error| vs fp64: split-TF32 {err3:.2e}, fp32 SGEMM {err_fp32:.2e}, single-pass TF32 {err1:.2e}")
    # The products are exact (hi*lo, lo*hi, hi*hi cover 22 mantissa bits); what remains is the tensor cores' accumulator, which
    # rounds toward zero at every k-step -- a bias proportional to the running sum.  With the small cross terms first along K the
    # large hi*hi part passes through 64 k-steps instead of 192: 1.8e-5 measured (7.6e-5 with hi*hi first; fp32 SGEMM 1.05e-5).
    assert err3 <= 3e-5, "the split GEMM must be as accurate as an fp32 GEMM to within a small factor"
    assert err3 <= 3 * err_fp32
    assert err1 > 10 * err3, "single-pass TF32 should be far worse (else this test does not exercise the split)"


@pytest.mark.parametrize("B,T,V,d,use_bias", [
    (8, 120, 5000, 512, True),    # BASELINE vocabulary and width; 960 rows = 7.5 row tiles, last vocabulary tile holds 136 columns
    (3, 50, 1000, 256, False),    # no bias, 150 rows (a partial second row tile), 4 vocabulary tiles = one per quarter
    (2, 40, 300, 64, True),       # 2 vocabulary tiles: two of the four quarters are empty; 2 k-blocks
    (5, 33, 260, 32, True),       # one k-block, V = 260: 4 valid columns in the second tile
    (2, 17, 6, 96, True),         # tiny vocabulary (blank inside): a single partial column group
    (3, 21, 40, 16, False),       # hidden size 16: half of the only k-block is zero padding
])
def test_log_posteriors_of_the_tcgen05_head_vs_fp64(B, T, V, d, use_bias):
    """What the scorer keeps -- padded log-posteriors and the blank column -- against an fp64 restatement of the reference's
    lm_head -> log_softmax -> padding (:39-46), ragged lengths incl. a zero-length utterance."""
    from huggingface_asr_b200.ctc_head import CTCHead

    g = torch.Generator().manual_seed(B * 1000 + V)
    weight = torch.randn(V, d, generator=g) / d ** 0.5 * 3.0
    bias = torch.randn(V, generator=g) if use_bias else None
    hidden = torch.randn(B, T, d, generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    if B > 2:
        lens[1] = 0
    z = hidden.double().reshape(-1, d) @ weight.double().t()
    if bias is not None:
        z = z + bias.double()
    ref = torch.log_softmax(z, -1).view(B, T, V)
    for b in range(B):
        ref[b, int(lens[b]):, :] = -1e10
        ref[b, int(lens[b]):, BLANK] = 0.0
    head = CTCHead(weight.cuda(), None if bias is None else bias.cuda(), implementation="tcgen05")
    x, blank_lp = head.log_posteriors(hidden.cuda(), lens.cuda(), BLANK)
    assert x.shape[:2] == (B, T) and x.shape[2] >= V and x.shape[2] % 64 == 0
    got = x[..., :V].cpu().double()
    err = float((got - ref).abs().max())
    zmax = float(z.abs().max())
    print(f"B={B} T={T} V={V} d={d}: max |log-posterior error| vs fp64 {err:.2e} (|logit| up to {zmax:.1f})")
    # the tensor-core accumulator rounds toward zero at every k-step: the error grows with |logit| (1.8e-5 at |z| ~ 12 through
    # 64 k-steps); these logits are drawn three times wider than a trained head's.  The path's tolerance is 1e-4.
    tol = 6e-5
    assert err <= tol
    assert float((blank_lp.cpu().double() - ref[..., BLANK]).abs().max()) <= tol
    logits = head(hidden.cuda())[..., :V].cpu().double().view(-1, V)
    assert float((logits - z).abs().max()) <= tol


def test_head_scaling_covers_the_fp32_range_fp16_lacks():
    """The tcgen05 head multiplies in fp16: rows of the hidden states whose magnitudes span 1e-5 .. 1e4 (and a weight far from 1)
    must come out as accurately as a row of ordinary size -- the error scale of a dot product is sum_k |h_k| |w_k|."""
    from huggingface_asr_b200.ctc_head import CTCHead

    g = torch.Generator().manual_seed(77)
    n, d, V = 300, 256, 520
    hidden = torch.randn(1, n, d, generator=g) * torch.logspace(-5, 4, n).view(1, n, 1)
    hidden[0, 7] = 0.0                                   # an all-zero row
    hidden[0, 9, :] = 0.0
    hidden[0, 9, 3] = 6.0e4                              # one huge element among zeros
    weight = torch.randn(V, d, generator=g) * 3.0e-3
    weight[5] *= 1.0e-4                                   # a vocabulary entry far below the tensor's scale
    bias = torch.randn(V, generator=g)
    ref = hidden.double().view(n, d) @ weight.double().t() + bias.double()
    scale = hidden.double().abs().view(n, d) @ weight.double().abs().t() + bias.double().abs()
    out = CTCHead(weight.cuda(), bias.cuda(), implementation="tcgen05")(hidden.cuda()).cpu().double().view(n, V)
    assert torch.isfinite(out).all()
    rel = ((out - ref).abs() / scale.clamp_min(1e-30)).max()
    print(f"max |error| / sum |h||w|: {float(rel):.2e}")
    assert float(rel) <= 1e-6


def test_both_head_implementations_feed_the_scorer_the_same_posteriors():
    from huggingface_asr_b200.ctc_head import CTCHead
    from huggingface_asr_b200.synthetic import make_encoder_hidden

    hidden, weight, bias, lens, _ = make_encoder_hidden(4, 96, 1000, 256, ragged=True, seed=9)
    outs = [CTCHead(weight.cuda(), bias.cuda(), implementation=i).log_posteriors(hidden.cuda(), lens.cuda(), BLANK) for i in ("tcgen05", "cublas")]
    assert float((outs[0][0][..., :1000] - outs[1][0][..., :1000]).abs().max()) <= 5e-5
    assert float((outs[0][1] - outs[1][1]).abs().max()) <= 5e-5


def test_processor_from_hidden_states_matches_the_oracle():
    from huggingface_asr_b200.beam_search import joint_beam_search, joint_beam_search_native
    from huggingface_asr_b200.ctc_head import CTCHead
    from huggingface_asr_b200.decoding.ctc_scorer import CTCRescorerLogitsProcessor
    from huggingface_asr_b200.synthetic import SyntheticDecoder, make_encoder_hidden
    from oracle import oracle as orc
    import parity

    B, W, T, V, d = 4, 10, 96, 1000, 256
    hidden, weight, bias, lens, transcripts = make_encoder_hidden(B, T, V, d, ragged=True, seed=6)
    logits_ref = (hidden.double().reshape(-1, d) @ weight.double().t() + bias.double()).view(B, T, V).float()
    head = CTCHead(weight.cuda(), bias.cuda())
    proc = CTCRescorerLogitsProcessor.from_encoder_hidden_states(hidden.cuda(), head, lens.cuda(), BLANK, EOS, 0, 0.3, W, -1, False, 1.0)
    cpu = orc.OracleCTCRescorerLogitsProcessor(logits_ref.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W)
    ids = torch.full((B * W, 1), BOS, dtype=torch.long)
    att = torch.log_softmax(torch.randn(B * W, V, generator=torch.Generator().manual_seed(2)), -1)
    parity.assert_parity(proc(ids.cuda(), att.cuda()), cpu(ids, att.clone()), "joint scores from hidden states")
    dec_g = SyntheticDecoder(transcripts, W, V, 40, seed=3, device="cuda")
    dec_c = SyntheticDecoder(transcripts, W, V, 40, seed=3, device="cpu")
    # the CPU and CUDA generators differ, so give both decoders the same noise
    dec_c.pool = [p.cpu() for p in dec_g.pool]
    out_g = joint_beam_search_native(CTCRescorerLogitsProcessor.from_encoder_hidden_states(hidden.cuda(), head, lens.cuda(), BLANK, EOS, 0,
                                                                                           0.3, W, -1, False, 1.0),
                                     dec_g, B, W, V, BOS, EOS, BLANK, max_length=40, device="cuda", done_check_lag=0)
    out_c = joint_beam_search(orc.OracleCTCRescorerLogitsProcessor(logits_ref.clone(), lens.clone(), BLANK, EOS, 0, 0.3, W), dec_c,
                              B, W, V, BOS, EOS, BLANK, max_length=40)
    assert out_g.steps == out_c.steps and (out_g.sequences.cpu() == out_c.sequences).all()
    assert (out_g.lengths.cpu() == torch.tensor([len(t) - 1 for t in transcripts])).all()
