"""All-in showdowns before the deal, two-card games - host side (CPU): tree construction and the float64 oracle.
The oracle's equity matrix is anchored on the reference's hand strengths (tests/golden/twocard_rows.npz holds the ranks of
lib_hand_eval.so): summing the golden brute-force showdown rows of the boards gives the all-in row (ValueFiller.py:160-175
generalised: the all-in terminal is a chance node whose children are showdowns)."""
import os

import numpy as np

import cfr2_numpy as o2
from gen_golden_twocard_common import make_reach
from pokerrl_b200.game.holdem_boards import BoardSpec
from twocard_common import fhp_tree, nl_flop_subgame, oracle_allin_equity, oracle_tree, random_board_spec

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "twocard_rows.npz"))


def test_short_stack_flop5_is_a_push_fold_tree():
    ft = fhp_tree(random_board_spec(8, 1), stack=300)
    assert ft.n_nodes == 5 and int((ft.kind == o2.KIND_SHOWDOWN_ALLIN).sum()) == 1 and int((ft.kind == o2.KIND_CHANCE).sum()) == 0
    assert ft.allin_spec is ft.board_spec
    order, _ = ft.work_order()
    lvl = ft.level_start
    for d in range(ft.n_levels):  # all-in terminals come last in every level
        k = ft.kind[order[lvl[d]:lvl[d + 1]]]
        assert np.all(np.diff(k.astype(int)) >= 0)


def test_oracle_allin_row_is_the_sum_of_the_golden_showdown_rows():
    """same opponent reach on 12 boards: weight-summed golden rows (float64 brute force on REFERENCE ranks) == E @ reach"""
    from pokerrl_b200.game import games
    rules = games.Flop5Holdem.RULES
    hc = np.asarray(rules.get_lut_holder().LUT_IDX_2_HOLE_CARDS).astype(np.int64)
    ids = np.array([0, 3, 17, 42, 77, 100, 121, 150, 160, 181, 190, 199])
    boards, ranks = GOLD["boards"][ids], GOLD["ranks"][ids]
    w = (np.arange(len(ids)) + 8.0) / 64.0  # exact in float32 (BoardSpec may store float32 probabilities)
    E = o2.allin_equity_matrix(ranks, w, hc, 52)
    assert np.abs(E + E.T).max() == 0.0  # zero-sum
    ro = make_reach(5, boards[:1], hc)[0].astype(np.float64)  # any reach row; hands on a board are masked per board by S_b
    want = np.zeros(1326)
    inc = np.zeros((1326, 52))
    inc[np.arange(1326), hc[:, 0]] = 1
    inc[np.arange(1326), hc[:, 1]] = 1
    compat = (inc @ inc.T) == 0
    for b in range(len(ids)):
        want += w[b] * (o2.sign_matrix(ranks[b], compat, ranks[b] < 0) @ ro)
    got = E @ ro
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # the evaluator pinned to lib_hand_eval.so gives the same matrix as the golden ranks
    spec = BoardSpec(boards, w, np.ones(len(ids)), None, "golden subset")
    assert np.array_equal(oracle_allin_equity(rules, spec), E)


def test_push_fold_oracle_converges():
    ft = fhp_tree(random_board_spec(24, 3), stack=300)
    t = oracle_tree(ft)
    cfr = o2.Oracle2CFR(t, "CFRPlus")
    e0 = cfr.exploitability_current()
    for _ in range(60):
        cfr.iteration()
    e1 = cfr.exploitability_average()
    assert e0 > 0 and 0 <= e1 < 0.1 * e0, (e0, e1)
    t.compute_ev()
    assert abs((t.ev[0] * t.reach[0]).sum()) < 1e-9  # zero-sum at the root (ValueFiller.py:98)


def test_nl_subgame_all_ins_on_the_flop_and_on_the_turn():
    """multi-street sub-game: one equity matrix per public board an all-in happens on; its completions carry the product of
    the remaining deal probabilities; the oracle stays zero-sum at the root and CFR+ converges"""
    ft = nl_flop_subgame()
    an = np.nonzero(ft.kind == o2.KIND_SHOWDOWN_ALLIN)[0]
    assert set(np.unique(ft.cdepth[an])) == {0, 1}
    comp = ft.allin_completions()
    assert set(comp) == set(int(b) for b in np.unique(ft.board[an]))
    flop = comp[int(ft.board[an[ft.cdepth[an] == 0][0]])]
    assert flop[0].shape == (12, 5) and abs(flop[1].sum() - 12 / (45 * 44)) < 1e-15  # 3 turn x 4 river cards of 45 x 44
    for key, (boards, w, _) in comp.items():
        if key != int(ft.board[an[ft.cdepth[an] == 0][0]]):
            assert boards.shape == (4, 5) and abs(w.sum() - 4 / 44) < 1e-15
    t = oracle_tree(ft)
    cfr = o2.Oracle2CFR(t, "CFRPlus")
    e0 = cfr.exploitability_current()
    for _ in range(40):
        cfr.iteration()
    assert 0 <= cfr.exploitability_average() < 0.2 * e0
    t.compute_ev()
    assert abs((t.ev[0] * t.reach[0]).sum()) < 1e-9 * np.abs(t.ev[0]).max()
