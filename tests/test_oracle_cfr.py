"""Pin the CPU oracle (oracle/cfr_numpy.py) against fixtures produced by the reference itself."""
import numpy as np
import pytest

import cfr_numpy as oc
from common import golden, make_flat_tree


def _dfs(ft, arr_flat):
    return arr_flat[ft.dfs_permutation()]


@pytest.mark.parametrize("name", ["StandardLeduc", "NLLeduc_POT", "NLLeduc_B2"])
def test_uniform_and_random_profile_values_bit_exact(name):
    ft = make_flat_tree(name)
    g = golden("values_%s.npz" % name)
    t = oc.OracleTree(ft)
    t.fill_uniform()
    expl = t.compute_ev()
    for k, mine in (("reach", t.reach), ("ev", t.ev), ("ev_br", t.ev_br)):
        assert np.array_equal(_dfs(ft, mine), g["uniform_" + k]), k
    assert np.array_equal(expl, g["uniform_root_exploitability"])
    # seeded random profile: load the reference's float64 strategies
    strat = g["random_strat"]  # [N, R] in DFS order, column of the parent's strategy
    perm = ft.dfs_permutation()
    inv = ft.dfs
    for n in t.decision_nodes():
        fc, A = ft.first_child[n], ft.n_children[n]
        t.strategy[n] = np.stack([strat[inv[fc + a]] for a in range(A)], axis=1)
    t.update_reach()
    expl = t.compute_ev()
    for k, mine in (("reach", t.reach), ("ev", t.ev), ("ev_br", t.ev_br)):
        assert np.array_equal(_dfs(ft, mine), g["random_" + k]), k
    assert np.array_equal(expl, g["random_root_exploitability"])


def test_b3_uniform_root():
    ft = make_flat_tree("NLLeduc_B3")
    g = golden("values_NLLeduc_B3.npz")
    t = oc.OracleTree(ft)
    t.fill_uniform()
    expl = t.compute_ev()
    assert np.array_equal(t.ev[0], g["uniform_root_ev"])
    assert np.array_equal(t.ev_br[0], g["uniform_root_ev_br"])
    assert np.array_equal(expl, g["uniform_root_exploitability"])


@pytest.mark.parametrize("algo,name,n_iters", [
    ("CFRPlus", "NLLeduc_POT", 31), ("CFRPlus", "StandardLeduc", 31),
    ("LinearCFR", "NLLeduc_POT", 31), ("LinearCFR", "StandardLeduc", 11),
    ("VanillaCFR", "NLLeduc_POT", 11), ("VanillaCFR", "StandardLeduc", 31),
])
def test_cfr_trajectory_bit_exact(algo, name, n_iters):
    ft = make_flat_tree(name)
    g = golden("cfr_%s_%s.npz" % (algo, name))
    cfr = oc.OracleCFR(ft, algo=algo)
    snaps = [t for t in (0, 1, 2, 3, 4, 5, 10, 11, 30, 31) if t <= n_iters]

    def check(t):
        if t > 0:
            assert np.array_equal(_dfs(ft, cfr.table(cfr.regret)), g["it%d_regret" % t], equal_nan=True), t
            assert np.array_equal(_dfs(ft, cfr.table(cfr.avg_strat)), g["it%d_avg" % t], equal_nan=True), t
        assert np.array_equal(_dfs(ft, cfr.table(cfr.tree.strategy)), g["it%d_strat" % t], equal_nan=True), t
        assert np.array_equal(_dfs(ft, cfr.tree.ev), g["it%d_ev" % t]), t
        assert np.array_equal(_dfs(ft, cfr.tree.ev_br), g["it%d_ev_br" % t]), t

    check(0)
    for t in range(1, n_iters + 1):
        cfr.iteration()
        if t in snaps:
            check(t)
    curr = np.array(cfr.curr_series)
    avg = np.array(cfr.avg_series)
    assert np.array_equal(curr, g["curr_series"][:len(curr)])
    assert np.array_equal(avg, g["avg_series"][:len(avg)])
