"""Seeded opponent reach rows shared by oracle/gen_golden_twocard.py (fixture generator) and the tests that consume
tests/golden/twocard_rows.npz.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

R = 1326


def make_reach(seed, boards, hand_cards):
    """float32 [n_boards, R]: skewed magnitudes, zero on hands holding a board card, every third board half-empty,
    every seventh concentrated on 40 hands"""
    rng = np.random.default_rng(seed)
    n = len(boards)
    r = (rng.random((n, R)) ** 3).astype(np.float32)
    for b in range(n):
        if b % 3 == 1:
            r[b, rng.random(R) < 0.5] = 0.0
        if b % 7 == 3:
            keep = rng.choice(R, 40, replace=False)
            m = np.zeros(R, bool)
            m[keep] = True
            r[b, ~m] = 0.0
        blocked = np.isin(hand_cards, boards[b]).any(axis=1)
        r[b, blocked] = 0.0
    return r
