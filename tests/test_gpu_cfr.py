"""GPU parity tests proper: the CUDA path (through the C ABI) against the reference's golden fixtures.

Bar (integer work: bit-exact; float work: the tolerance BASELINE.json states, 1e-6 relative): in practice every
float32 quantity of the one-card games is reproduced BIT-FOR-BIT, because the kernels evaluate in the reference's
dtypes and operation order — the tests therefore assert exact equality and would catch any
reordering.
"""
import numpy as np
import pytest
import torch

from common import golden, make_flat_tree

pytestmark = pytest.mark.gpu


def _node_vec(ft, t):  # torch [2, N, ld] -> numpy [N, 2, R] in DFS order
    a = t.cpu().numpy()[:, :, :ft.R].transpose(1, 0, 2)
    return a[ft.dfs_permutation()]


def _table(ft, t):  # torch [n_slots, ld] -> numpy [N, R] (NaN where the node has no slot) in DFS order
    a = t.cpu().numpy()[:, :ft.R].astype(np.float64)
    out = np.full((ft.n_nodes, ft.R), np.nan)
    m = ft.slot >= 0
    out[m] = a[ft.slot[m]]
    return out[ft.dfs_permutation()]


@pytest.mark.parametrize("name", ["StandardLeduc", "NLLeduc_POT", "NLLeduc_B2"])
def test_uniform_and_random_profile_bit_exact(name):
    from pokerrl_b200 import _native as nat
    from pokerrl_b200.solver import CFRSolver
    ft = make_flat_tree(name)
    g = golden("values_%s.npz" % name)
    s = CFRSolver(ft, "CFRPlus", avg_f64=True)
    s.exploitability_current()
    for k, t in (("reach", s.bufs.reach), ("ev", s.bufs.ev), ("ev_br", s.bufs.ev_br)):
        assert np.array_equal(_node_vec(ft, t), g["uniform_" + k]), k
    assert np.array_equal(s.ops.root_exploitability(), g["uniform_root_exploitability"])
    # seeded random float64 profile of the reference -> double table, evaluated like an average strategy
    strat = g["random_strat"]  # [N, R] DFS order
    tab = np.zeros((ft.n_slots, ft.R))
    m = ft.slot >= 0
    tab[ft.slot[m]] = strat[ft.dfs[m]]
    s.bufs.avg.copy_(torch.from_numpy(tab))
    modes = [nat.STRAT_AVG_F64, nat.STRAT_AVG_F64]
    s.ops.reach_pass(modes)
    s.ops.value_pass(modes, 3, True)
    for k, t in (("reach", s.bufs.reach), ("ev", s.bufs.ev), ("ev_br", s.bufs.ev_br)):
        assert np.array_equal(_node_vec(ft, t), g["random_" + k]), k
    assert np.array_equal(s.ops.root_exploitability(), g["random_root_exploitability"])


def test_b3_uniform_root():
    from pokerrl_b200.solver import CFRSolver
    ft = make_flat_tree("NLLeduc_B3")
    g = golden("values_NLLeduc_B3.npz")
    s = CFRSolver(ft, "CFRPlus")
    s.exploitability_current()
    assert np.array_equal(s.bufs.ev[:, 0, :ft.R].cpu().numpy(), g["uniform_root_ev"])
    assert np.array_equal(s.bufs.ev_br[:, 0, :ft.R].cpu().numpy(), g["uniform_root_ev_br"])
    assert np.array_equal(s.ops.root_exploitability(), g["uniform_root_exploitability"])


@pytest.mark.parametrize("algo,name", [
    ("CFRPlus", "NLLeduc_POT"), ("CFRPlus", "StandardLeduc"),
    ("LinearCFR", "NLLeduc_POT"), ("LinearCFR", "StandardLeduc"),
    ("VanillaCFR", "NLLeduc_POT"), ("VanillaCFR", "StandardLeduc"),
])
@pytest.mark.parametrize("persistent", [True, False])
def test_cfr_trajectory_bit_exact(algo, name, persistent):
    from pokerrl_b200.solver import CFRSolver
    ft = make_flat_tree(name)
    g = golden("cfr_%s_%s.npz" % (algo, name))
    s = CFRSolver(ft, algo, avg_f64=True, persistent=persistent)
    n_iters = 40 if algo != "CFRPlus" else (150 if name == "NLLeduc_POT" else 60)
    curr, avg = [(0, s.exploitability_current())], []
    snaps = (1, 2, 3, 4, 5, 10, 11, 30, 31)
    for t in range(1, n_iters + 1):
        s.iteration()
        curr.append((t, s.exploitability_current()))
        if t in snaps:
            assert np.array_equal(_table(ft, s.bufs.regret), g["it%d_regret" % t], equal_nan=True), t
            assert np.array_equal(_table(ft, s.bufs.strat), g["it%d_strat" % t], equal_nan=True), t
            for k, buf in (("reach", s.bufs.reach), ("ev", s.bufs.ev), ("ev_br", s.bufs.ev_br)):
                assert np.array_equal(_node_vec(ft, buf), g["it%d_%s" % (t, k)]), (t, k)
            if algo == "CFRPlus":
                assert np.array_equal(_table(ft, s.bufs.avg), g["it%d_avg" % t], equal_nan=True), t
            else:
                assert np.array_equal(_table(ft, s.bufs.avg), g["it%d_avg_sum" % t], equal_nan=True), t
        avg.append((t, s.exploitability_average()))
    assert np.array_equal(np.array(curr), g["curr_series"]), "current-strategy exploitability series"
    assert np.array_equal(np.array(avg), g["avg_series"]), "average-strategy exploitability series"


def test_cfr_plus_float32_average_within_tolerance():
    """The float32 average table (reference semantics under numpy<2, and the fast default) stays within the
    1e-6 relative tolerance of the float64-average reference run."""
    from pokerrl_b200.solver import CFRSolver
    ft = make_flat_tree("NLLeduc_POT")
    g = golden("cfr_CFRPlus_NLLeduc_POT.npz")
    s = CFRSolver(ft, "CFRPlus", avg_f64=False)
    for t in range(1, 61):
        s.iteration()
        a = s.exploitability_average()
        ref = g["avg_series"][t - 1, 1]
        assert abs(a - ref) <= 1e-6 * abs(ref), (t, a, ref)


def test_persistent_multi_iteration_launch_matches_golden():
    """20 iterations inside ONE cooperative launch (bench cadence) land on the reference's iteration-20/40 numbers."""
    from pokerrl_b200.solver import CFRSolver
    ft = make_flat_tree("NLLeduc_POT")
    g = golden("cfr_CFRPlus_NLLeduc_POT.npz")
    s = CFRSolver(ft, "CFRPlus", avg_f64=True)
    for t in (20, 40, 60):
        s.iteration(20)
        assert s.exploitability_current() == g["curr_series"][t, 1]
        assert s.exploitability_average() == g["avg_series"][t - 1, 1]


def test_local_br_master_with_tabular_agent_matches_internal_average_evaluation():
    """LocalBRMaster(fill_with_agent_policy -> reach -> value+BR) on the tabular agent == the solver's own average-
    strategy exploitability (LocalBRMaster.py:41-80 vs _CFRBase.py:218-262); also the pickle round trip."""
    import os
    import tempfile
    from pokerrl_b200.cfr.CFRPlus import CFRPlus
    from pokerrl_b200.cfr.TabularCFREvalAgent import TabularCFREvalAgent, average_strategy_table
    from pokerrl_b200.eval.br.LocalBRMaster import LocalBRMaster
    from pokerrl_b200.game import bet_sets
    from pokerrl_b200.game.games import DiscretizedNLLeduc
    from pokerrl_b200.rl.base_cls.TrainingProfileBase import TrainingProfileBase
    from pokerrl_b200.rl.base_cls.workers.ChiefBase import ChiefBase
    chief = ChiefBase(t_prof=None)
    cfr = CFRPlus(name="k", chief_handle=chief, game_cls=DiscretizedNLLeduc, agent_bet_set=bet_sets.POT_ONLY, delay=0)
    for _ in range(12):
        cfr.iteration()
    ref = chief.get_experiments()["k_Avg_total_S20000_CFRp_delay0"]["Evaluation/MBB_per_G"][-1][1]
    t_prof = TrainingProfileBase("k", DiscretizedNLLeduc, bet_sets.POT_ONLY)
    br = LocalBRMaster(t_prof=t_prof, chief_handle=chief, eval_agent_cls=TabularCFREvalAgent)
    br.eval_agent.update_weights(average_strategy_table(cfr.solvers[0]))
    br.evaluate(iter_nr=12)
    got = chief.get_experiments()["k AVG_stack_20000: BR Total"]["Evaluation/MBB_per_G"][-1]
    assert got[0] == 12 and got[1] == ref
    with tempfile.TemporaryDirectory() as d:
        br.eval_agent.store_to_disk(d, "agent")
        again = TabularCFREvalAgent.load_from_disk(os.path.join(d, "agent.pkl"))
        assert np.array_equal(again._table, br.eval_agent._table)


def test_checkpoint_resume_continues_the_same_trajectory(tmp_path):
    from pokerrl_b200.cfr.LinearCFR import LinearCFR
    from pokerrl_b200.game import bet_sets
    from pokerrl_b200.game.games import StandardLeduc
    from pokerrl_b200.rl.base_cls.workers.ChiefBase import ChiefBase
    g = golden("cfr_LinearCFR_StandardLeduc.npz")
    a = LinearCFR(name="a", chief_handle=ChiefBase(None), game_cls=StandardLeduc, agent_bet_set=bet_sets.POT_ONLY, avg_f64=False)
    for _ in range(5):
        a.iteration()
    a.checkpoint(str(tmp_path / "ck.pt"))
    b = LinearCFR(name="b", chief_handle=ChiefBase(None), game_cls=StandardLeduc, agent_bet_set=bet_sets.POT_ONLY)
    b.load_checkpoint(str(tmp_path / "ck.pt"))
    for _ in range(5):
        b.iteration()
    assert b.iter_counter == 10
    assert b.solvers[0].exploitability_current() == g["curr_series"][10, 1]
    assert b.solvers[0].exploitability_average() == g["avg_series"][9, 1]


def test_batched_fill_with_agent_policy_equals_the_per_node_loop():
    """N2: the agent answers for all decision nodes at once and one device gather fills the strategy table;
    same table, same exploitability as the reference-style node-by-node query loop (StrategyFiller.py:88-116) - NL-Leduc and
    a 32-board Flop5Holdem tree."""
    import torch
    from pokerrl_b200.cfr.TabularCFREvalAgent import TabularCFREvalAgent
    from pokerrl_b200.game.PublicTree import PublicTree
    from pokerrl_b200.game.flat_tree import FlatTree
    from pokerrl_b200.game.games import DiscretizedNLLeduc, Flop5Holdem
    from pokerrl_b200.game.wrappers import HistoryEnvBuilder
    from pokerrl_b200.rl.base_cls.TrainingProfileBase import TrainingProfileBase
    from pokerrl_b200.solver import CFRSolver
    from pokerrl_b200.game import bet_sets
    import os
    from twocard_common import random_board_spec
    for game, bet_set, spec in ((DiscretizedNLLeduc, bet_sets.POT_ONLY, None), (Flop5Holdem, [1.0], random_board_spec(32, 4))):
        args = game.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[game.DEFAULT_STACK_SIZE] * 2,
                             bet_sizes_list_as_frac_of_pot=list(bet_set))
        ft = FlatTree(game, args, board_spec=spec)
        os.environ["PRL_ENGINE"] = "levels"
        try:
            s = CFRSolver(ft, "CFRPlus")
            s.iteration(5)
        finally:
            os.environ.pop("PRL_ENGINE", None)
        t_prof = TrainingProfileBase(name="n2", game_cls=game, agent_bet_set=list(bet_set), eval_stack_sizes=None)
        agent = TabularCFREvalAgent(t_prof=t_prof)
        from pokerrl_b200.cfr.TabularCFREvalAgent import average_strategy_table, tree_fingerprint
        agent.update_weights((average_strategy_table(s), tree_fingerprint(ft)))
        bldr = HistoryEnvBuilder(env_cls=game, env_args=args)
        trees = []
        for batched in (True, False):
            tree = PublicTree(env_bldr=bldr, stack_size=args.starting_stack_sizes_list, stop_at_street=None, board_spec=spec)
            tree.build_tree()
            if not batched:
                agent.get_a_probs_for_public_tree = lambda tree: None
            tree.fill_with_agent_policy(agent=agent)
            tree.compute_ev()
            trees.append(tree)
        assert torch.equal(trees[0].bufs.strat, trees[1].bufs.strat)
        assert np.array_equal(trees[0].root.exploitability, trees[1].root.exploitability)
        assert abs(float(np.mean(trees[0].root.exploitability)) * game.EV_NORMALIZER - s.exploitability_average()) <= \
            1e-5 * s.exploitability_average()
