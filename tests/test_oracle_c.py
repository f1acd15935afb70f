"""Pin the C oracle (oracle/cfr_oracle.c) against the reference's golden fixtures, bit for bit."""
import numpy as np
import pytest

import cfr_c
from common import golden, make_flat_tree


def _node_vec(ft, a):
    return a.transpose(1, 0, 2)[ft.dfs_permutation()]


def _table(ft, a):
    out = np.full((ft.n_nodes, ft.R), np.nan)
    m = ft.slot >= 0
    out[m] = a[ft.slot[m]].astype(np.float64)
    return out[ft.dfs_permutation()]


@pytest.mark.parametrize("algo,name", [
    ("CFRPlus", "NLLeduc_POT"), ("CFRPlus", "StandardLeduc"), ("LinearCFR", "NLLeduc_POT"),
    ("LinearCFR", "StandardLeduc"), ("VanillaCFR", "NLLeduc_POT"), ("VanillaCFR", "StandardLeduc"),
])
def test_c_oracle_trajectory_bit_exact(algo, name):
    ft = make_flat_tree(name)
    g = golden("cfr_%s_%s.npz" % (algo, name))
    s = cfr_c.OracleCSolver(ft, algo, avg_f64=True)
    n_iters = len(g["curr_series"]) - 1
    curr, avg = [(0, s.exploitability_current())], []
    assert np.array_equal(_node_vec(ft, s.ev), g["it0_ev"])
    assert np.array_equal(_node_vec(ft, s.ev_br), g["it0_ev_br"])
    for t in range(1, n_iters + 1):
        s.iteration()
        curr.append((t, s.exploitability_current()))
        if t in (1, 2, 3, 10, 31):
            assert np.array_equal(_table(ft, s.regret), g["it%d_regret" % t], equal_nan=True), t
            assert np.array_equal(_table(ft, s.strat), g["it%d_strat" % t], equal_nan=True), t
            assert np.array_equal(_node_vec(ft, s.reach), g["it%d_reach" % t]), t
            key = "it%d_avg" % t if algo == "CFRPlus" else "it%d_avg_sum" % t
            assert np.array_equal(_table(ft, s.avg), g[key], equal_nan=True), t
        avg.append((t, s.exploitability_average()))
    assert np.array_equal(np.array(curr), g["curr_series"])
    assert np.array_equal(np.array(avg), g["avg_series"])


def test_c_oracle_b3_uniform_root():
    ft = make_flat_tree("NLLeduc_B3")
    g = golden("values_NLLeduc_B3.npz")
    s = cfr_c.OracleCSolver(ft, "CFRPlus")
    s.exploitability_current()
    assert np.array_equal(s.ev[:, 0], g["uniform_root_ev"])
    assert np.array_equal(s.ev_br[:, 0], g["uniform_root_ev_br"])
