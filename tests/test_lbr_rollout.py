"""LBR roll-out: the golden equities come from the REFERENCE's _LBRRolloutManager (oracle/gen_golden_lbr.py).
CPU part: a numpy restatement of the kernel's arithmetic reproduces them (so the statement of the algorithm is pinned without
a GPU); GPU part: csrc/lbr_rollout.cu through the C ABI."""
import os
from itertools import combinations

import numpy as np
import pytest

from twocard_common import oracle_ranks

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lbr_rollouts.npz"))
TOL = 2e-5  # the reference accumulates in float32 (numpy scalar promotion), the kernel in float64


def _restate(hand, board, k, rng, quirk):
    """quirk: ranks of completion 0 for every completion - what the reference does (its `_i` is never advanced,
    LocalLBRWorker.py:468-512)"""
    hc = np.array(list(combinations(range(52), 2)))
    acp = np.zeros(52, np.float32)
    for c in range(52):
        acp[c] = rng[(hc == c).any(axis=1)].sum(dtype=np.float32)
    cp = (1.0 - acp).astype(np.float64)
    cp[list(hand) + list(board[:k])] = 0.0
    if cp.sum() > 0:
        cp /= cp.sum()
    poss = [c for c in range(52) if c not in hand and c not in board[:k]]
    combos = list(combinations(poss, 5 - k))
    full = np.array([list(board[:k]) + list(c) for c in combos], np.int8)
    ranks = oracle_ranks(full) if len(full) else None
    lbr_idx = int(np.nonzero((hc == sorted(hand)).all(axis=1))[0][0])
    tot = 0.0
    for t, c in enumerate(combos):
        reach, left = 1.0, 1.0
        for x in c:
            reach *= cp[x] / left if left > 0 else 0.0
            left -= cp[x]
        rk = ranks[0] if quirk else ranks[t]
        live = ranks[t] >= 0
        w = np.where(live, rng.astype(np.float64), 0.0)
        z = w.sum()
        if z > 0:
            eq = (w[rk < rk[lbr_idx]].sum() + 0.5 * w[rk == rk[lbr_idx]].sum()) / z
        else:
            eq = ((rk < rk[lbr_idx]).sum() + 0.5 * (rk == rk[lbr_idx]).sum()) / 1326.0
        tot += eq * reach
    return tot * float(np.prod(np.arange(1, 5 - k + 1)))


def test_numpy_restatement_reproduces_the_reference_rollouts():
    for q in range(len(GOLD["equity"])):
        if GOLD["n_dealt"][q] == 3 and q > 0:
            continue  # one flop query is enough on the CPU (1081 boards each)
        got = _restate(GOLD["hands"][q].tolist(), GOLD["boards"][q].tolist(), int(GOLD["n_dealt"][q]), GOLD["ranges"][q], True)
        assert abs(got - GOLD["equity"][q]) <= TOL * max(GOLD["equity"][q], 1e-3), (q, got, GOLD["equity"][q])


@pytest.mark.gpu
def test_gpu_rollouts_match_the_reference():
    from pokerrl_b200.eval.lbr.rollout import LBRRolloutManager, lbr_checkdown_equity
    worst = 0.0
    for k in (3, 4, 5):
        sel = np.nonzero(GOLD["n_dealt"] == k)[0]
        out = lbr_checkdown_equity(GOLD["hands"][sel], GOLD["boards"][sel], k, GOLD["ranges"][sel],
                                   reference_board_counter_quirk=True).cpu().numpy()
        err = np.abs(out - GOLD["equity"][sel]) / np.maximum(GOLD["equity"][sel], 1e-3)
        worst = max(worst, float(err.max()))
        assert np.all(err <= TOL), (k, out, GOLD["equity"][sel])
        # the product's default ranks every completion on its own cards: against the numpy restatement without the quirk
        q = int(sel[1])
        ref = _restate(GOLD["hands"][q].tolist(), GOLD["boards"][q].tolist(), k, GOLD["ranges"][q], False)
        got = float(lbr_checkdown_equity(GOLD["hands"][q:q + 1], GOLD["boards"][q:q + 1], k, GOLD["ranges"][q:q + 1])[0].item())
        assert abs(got - ref) <= 1e-6, (k, got, ref)
    print("LBR roll-out vs reference: worst relative error %.1e" % worst)

    class Env:  # the manager reads the public board from the env it is given (LocalLBRWorker.py:389-390)
        pass
    from pokerrl_b200.game import bet_sets, games
    from pokerrl_b200.game.wrappers import HistoryEnvBuilder
    g = games.DiscretizedNLHoldem
    bldr = HistoryEnvBuilder(env_cls=g, env_args=g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000],
                                                            bet_sizes_list_as_frac_of_pot=bet_sets.B_2))
    q = int(np.nonzero(GOLD["n_dealt"] == 4)[0][0])
    env = Env()
    env.board = bldr.lut_holder.get_2d_cards(GOLD["boards"][q])
    m = LBRRolloutManager(t_prof=None, env_bldr=bldr, env=env, lbr_hand_2d=bldr.lut_holder.get_2d_cards(GOLD["hands"][q]),
                          reference_board_counter_quirk=True)
    assert abs(m.get_lbr_checkdown_equity(GOLD["ranges"][q]) - GOLD["equity"][q]) <= TOL
