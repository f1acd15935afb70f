"""Pins the two-card oracles' terminal formulas at R = 1326 on tests/golden/twocard_rows.npz: brute-force O(R^2) float64
rows of ValueFiller.py:103-158 (generalised) computed with hand strengths from the REFERENCE's
lib_hand_eval.so on 200 boards incl. tie-heavy ones (oracle/gen_golden_twocard.py)."""
import os

import numpy as np

import cfr2_c
import cfr2_numpy as o2
from gen_golden_twocard_common import make_reach
from twocard_common import oracle_ranks

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "twocard_rows.npz"))


def _hand_cards():
    from pokerrl_b200.game.games import Flop5Holdem
    return np.asarray(Flop5Holdem.RULES.get_lut_holder().LUT_IDX_2_HOLE_CARDS).astype(np.int64)


def test_reach_rows_regenerate():
    hc = _hand_cards()
    reach = make_reach(int(GOLD["seed"]), GOLD["boards"], hc)
    assert np.array_equal(reach.astype(np.float64).sum(axis=1), GOLD["reach_sum"])


def test_c_hand_eval_oracle_equals_reference_ranks_on_these_boards():
    assert np.array_equal(oracle_ranks(GOLD["boards"]), GOLD["ranks"])


def test_numpy_and_c_oracle_rows_equal_brute_force():
    hc = _hand_cards()
    boards, ranks, K = GOLD["boards"], GOLD["ranks"], float(GOLD["eq_const"])
    reach = make_reach(int(GOLD["seed"]), boards, hc).astype(np.float64)
    R = 1326
    inc = np.zeros((R, 52))
    inc[np.arange(R), hc[:, 0]] = 1
    inc[np.arange(R), hc[:, 1]] = 1
    compat = (inc @ inc.T) == 0
    worst = 0.0
    for b in range(len(boards)):
        blocked = ranks[b] < 0
        sd = K * (o2.sign_matrix(ranks[b], compat, blocked) @ reach[b])
        fo = K * o2.fold_row(inc, reach[b])
        fo[blocked] = 0.0
        csd, cfo = cfr2_c.terminal_rows(hc, ranks[b], reach[b], K)
        scale = max(np.abs(GOLD["showdown"][b]).max(), np.abs(GOLD["fold"][b]).max(), 1e-300)
        for name, got, ref in (("numpy showdown", sd, GOLD["showdown"][b]), ("numpy fold", fo, GOLD["fold"][b]),
                               ("C showdown", csd, GOLD["showdown"][b]), ("C fold", cfo, GOLD["fold"][b])):
            err = np.abs(got - ref).max() / scale
            worst = max(worst, err)
            assert err < 1e-12, (b, name, err)
    print("worst relative deviation from the brute-force rows: %.2e" % worst)
