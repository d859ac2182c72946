"""The CPU oracle against the UNMODIFIED reference scorer imported from /root/reference, on randomly drawn shapes.

The committed fixtures (tests/golden/) pin the oracle on fixed seeds; this test widens the net where the reference is
present (the dev container): random B, W, T, V, ragged lengths, peaky and flat posteriors, several decode steps with beam
reordering, every step compared tensor by tensor.  On a box without /root/reference (the GPU box) it is skipped: nothing
there may read the reference.  CPU only.
"""
import importlib.util
import os

import pytest
import torch

import parity
from huggingface_asr_b200.synthetic import make_attention_scores, make_encoder_logits
from oracle import oracle as orc

REF = "/root/reference/src/decoding/ctc_scorer.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the reference is not on this box")


def _reference():
    spec = importlib.util.spec_from_file_location("_reference_ctc_scorer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("seed", range(8))
def test_oracle_processor_equals_the_live_reference(seed):
    ref_mod = _reference()
    g = torch.Generator().manual_seed(9000 + seed)
    B = int(torch.randint(1, 4, (1,), generator=g))
    W = int(torch.randint(1, 7, (1,), generator=g))
    T = int(torch.randint(6, 40, (1,), generator=g))
    V = int(torch.randint(8, 70, (1,), generator=g))
    kind = "peaky" if seed % 2 == 0 else "flat"
    w = [0.3, 0.5, 1.0][seed % 3]
    logits, lens, _ = make_encoder_logits(B, T, V, kind, True, seed=500 + seed)
    ref = ref_mod.CTCRescorerLogitsProcessor(logits.clone(), lens.clone(), 3, 1, 0, w, W, -1, False, 1.0)
    ora = orc.OracleCTCRescorerLogitsProcessor(logits.clone(), lens.clone(), 3, 1, 0, w, W, -1, False, 1.0)
    ref64 = ref_mod.CTCRescorerLogitsProcessor(logits.double(), lens.clone(), 3, 1, 0, w, W, -1, False, 1.0)
    ids = torch.zeros((B * W, 1), dtype=torch.long)
    beam_scores = torch.zeros(B, W)
    beam_scores[:, 1:] = -1e9
    for n in range(min(6, T - 1)):
        att = make_attention_scores(B * W, V, n, seed=700 + seed, scale=0.5)
        a_ref, a_ora = att.clone(), att.clone()
        out_r = ref(ids, a_ref)
        out_o = ora(ids, a_ora)
        out_64 = ref64(ids, att.double())
        what = f"seed {seed} (B={B} W={W} T={T} V={V} {kind}) step {n}"
        assert torch.equal(a_ref, a_ora), f"{what}: in-place pad column of the caller's scores"
        parity.assert_parity(out_o, out_r, f"{what} joint scores", ref64=out_64)
        parity.assert_parity(ora.ctc_states[1], ref.ctc_states[1], f"{what} log_psi", ref64=ref64.ctc_states[1])
        parity.assert_parity(ora.ctc_states[0], ref.ctc_states[0], f"{what} r", ref64=ref64.ctc_states[0])
        cand = (out_r + beam_scores.view(-1, 1)).view(B, W * V)
        top, idx = cand.topk(W, dim=1)
        src, tok = idx // V, idx % V
        ids = torch.cat([ids[(src + (torch.arange(B) * W).view(B, 1)).view(-1)], tok.view(-1, 1)], dim=1)
        beam_scores = top
