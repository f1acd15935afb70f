"""The two-card-generalised float64 oracle (oracle/cfr2_numpy.py) reduces to the reference-pinned one-card oracle."""
import numpy as np
import pytest

import cfr2_numpy as o2
import cfr_numpy as o1
from common import golden, make_flat_tree


def leduc_oracle2(ft):
    rules = ft.rules
    R = ft.R
    bc = ft.board_cards()
    ranks = np.full((bc.shape[0], R), -1, np.int32)
    for b in range(bc.shape[0]):
        if bc[b, 0] >= 0:
            ranks[b] = [o1.leduc_hand_rank(h, int(bc[b, 0]), rules.N_SUITS, rules.PAIR_BONUS) for h in range(R)]
    nb = bc.shape[0]
    return o2.Oracle2Tree(ft, np.arange(R).reshape(-1, 1), ranks, np.full(nb, 1.0 / (rules.N_CARDS_IN_DECK - 2)),
                          np.ones(nb))


def test_reduces_to_reference_values_on_standard_leduc():
    ft = make_flat_tree("StandardLeduc")
    g = golden("values_StandardLeduc.npz")
    t = leduc_oracle2(ft)
    t.fill_uniform()
    expl = t.compute_ev()
    perm = ft.dfs_permutation()
    for k, mine in (("reach", t.reach), ("ev", t.ev), ("ev_br", t.ev_br)):
        ref = g["uniform_" + k].astype(np.float64)
        assert np.allclose(mine[perm], ref, rtol=2e-6, atol=2e-6 * np.abs(ref).max()), k
    assert np.allclose(expl, g["uniform_root_exploitability"], rtol=1e-6)


@pytest.mark.parametrize("algo", ["CFRPlus", "LinearCFR", "VanillaCFR"])
def test_first_iterations_match_reference_series(algo):
    """iterations 0-1 agree with the reference to float32 round-off (later ones are decided by its float32 noise,
    SURVEY.md appendix C, so only the envelope is comparable)."""
    ft = make_flat_tree("StandardLeduc")
    g = golden("cfr_%s_StandardLeduc.npz" % algo)
    c = o2.Oracle2CFR(leduc_oracle2(ft), algo, ev_normalizer=ft.game_cls.EV_NORMALIZER)
    assert np.isclose(c.exploitability_current(), g["curr_series"][0, 1], rtol=1e-6)
    c.iteration()
    assert np.isclose(c.exploitability_current(), g["curr_series"][1, 1], rtol=1e-6)
    assert np.isclose(c.exploitability_average(), g["avg_series"][0, 1], rtol=1e-6)
    for _ in range(30):
        c.iteration()
    ref = g["avg_series"][30, 1]
    assert abs(c.exploitability_average() - ref) < 0.1 * ref  # same convergence envelope
