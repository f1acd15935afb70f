"""Bet sets (pot fractions) for discretized no-limit games; numeric contents identical to the reference's
`PokerRL/game/bet_sets.py:11-143` (they are plain data a user passes as `agent_bet_set`)."""

_ALLIN = 100000.0


def _with_allin(*fracs):
    return list(fracs) + [_ALLIN]


ALL_IN_ONLY = _with_allin()
POT_ONLY = [1.0]
B_2 = _with_allin(1.0)
B_3 = _with_allin(0.5, 1.0)
B_4 = _with_allin(0.5, 1.0, 2.0)
B_5_SMALL = _with_allin(0.5, 1.0, 2.0, 2.3)
B_5 = _with_allin(0.5, 0.75, 1.0, 2.0)
B_8 = _with_allin(0.5, 0.7, 1.25, 1.5, 2.0, 2.5)
B_16 = _with_allin(0.10, 0.25, 0.40, 0.55, 0.75, 0.90, 1.10, 1.30, 1.50, 1.80, 2.25, 3.30, 4.50, 6.00, 8.00)
B_21 = _with_allin(0.10, 0.22, 0.33, 0.44, 0.55, 0.75, 0.88, 1.00, 1.10, 1.25, 1.40, 1.60, 1.80, 2.00, 2.25, 2.75,
                   3.75, 4.00, 5.50, 8.00)
PL_6 = [0.1, 0.22, 0.3, 0.50, 0.73, 1.0]
PL_10 = [0.1, 0.15, 0.22, 0.3, 0.39, 0.50, 0.61, 0.73, 0.86, 1.0]
OFF_TREE_1 = _with_allin(0.7)
OFF_TREE_5 = _with_allin(0.38, 0.63, 0.93, 1.73)
OFF_TREE_11 = _with_allin(0.20, 0.42, 0.52, 0.86, 1.23, 1.65, 2.05, 3.40, 5.00, 7.0)
