"""Pins the CPU oracle (oracle/ctc_prefix_oracle.c) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import parity
from oracle import oracle as orc


class OracleBackend(parity.Backend):
    device = "cpu"

    def make_scorer(self, x_logp, lens, blank, eos, margin=0):
        return orc.OracleCTCPrefixScore(x_logp, lens, blank, eos, margin)

    def make_processor(self, logits, lens, pad, eos, margin, w, W, space=-1, trick=False, trick_w=1.0):
        return orc.OracleCTCRescorerLogitsProcessor(logits, lens, pad, eos, margin, w, W, space, trick, trick_w)


BE = OracleBackend()
STEP_CASES = ["steps_peaky_w3", "steps_peaky_ragged_w10", "steps_flat_w1", "steps_flat_w20", "steps_peaky_w5_v129",
              "steps_forced_pad", "steps_trick"]


@pytest.mark.parametrize("name", STEP_CASES)
def test_oracle_steps(name):
    worst = parity.replay_steps(BE, name)
    print(name, worst)


def test_oracle_partial_scoring():
    parity.replay_partial(BE)


def test_oracle_select_general():
    parity.replay_select_general(BE)


def test_oracle_edges():
    parity.replay_edges(BE)


@pytest.mark.parametrize("name", parity.WINDOW_CASES)
def test_oracle_attention_window(name):
    """margin > 0 with attention weights: the frame window of ctc_scorer.py:127-136."""
    print(name, parity.replay_window(BE, name))


def test_oracle_extend_prob_and_state():
    parity.replay_extend(BE)


def test_oracle_decode_1best():
    parity.replay_decode(BE)


@pytest.mark.parametrize("name", parity.PREBEAM_CASES)
def test_oracle_prebeam_policy_vs_reference_scorer(name):
    """ESPnet's pre-beam policy around the scorer (scoring_ids = top-S decoder tokens, hyp*V+tok state selection)."""
    def mk(logits, lens, w, W, S, use_beam_idx):
        return orc.OracleCTCRescorerLogitsProcessor(logits, lens, 3, 1, 0, w, W, pre_beam_size=S, use_beam_idx=use_beam_idx)

    print(name, parity.replay_prebeam(mk, "cpu", name))


def test_oracle_padded_posteriors_match_reference():
    g = parity.load("steps_peaky_ragged_w10")
    x = orc.log_softmax(g["logits"])
    orc.pad(x, g["lens"], 3)
    parity.assert_parity(x, g["x_padded"], "padded log-posteriors", atol=2e-6, rtol=0)


def test_oracle_fp64_agrees_with_reference_fp64():
    g = parity.load("steps_flat_w20")
    x = orc.log_softmax(g["logits"].astype(np.float64), 64)
    orc.pad(x, g["lens"], 3, 64)
    W = int(g["W"])
    r0 = orc.init_state(x, 3, W, 64)
    ts, r, log_psi, _, _ = orc.score(x, 3, r0, None, np.zeros(W, dtype=np.int64), 0, W, None, 64)
    parity.assert_parity(ts, g["token_scores_0_f64"], "fp64 token_scores", atol=1e-9, rtol=1e-12)
    parity.assert_parity(r, g["r_0_f64"], "fp64 r", atol=1e-9, rtol=1e-12)
