"""Single-table PokerEnv facade (pokerrl_b200/game/poker_env.py) and the drop-in acceptance run of
examples/run_cfrp_example.py (BASELINE.json configs[0]) through the PokerRL compat namespace."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["DiscretizedNLHoldem_B5", "LimitHoldem", "DiscretizedNLLeduc_B3"])
def test_reset_step_replays_reference_trajectories(name, golden_dir):
    """reset(deck_state_dict) / step / get_legal_actions on decks + actions recorded from the REFERENCE env
    (oracle/gen_golden_env.py): identical observations, rewards, done flags and legal action lists; state_dict round trip"""
    from pokerrl_b200.game import bet_sets, games
    from pokerrl_b200.game.poker_env import PokerEnv
    g = np.load(os.path.join(golden_dir, "env_%s.npz" % name))
    game = getattr(games, str(g["game"]))
    stack = int(g["stack"])
    args = game.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[stack, stack],
                         bet_sizes_list_as_frac_of_pot=list(getattr(bet_sets, str(g["bet_set"]))))
    env = PokerEnv(game, args, is_evaluating=True)
    lut, nh = env.lut_holder, game.RULES.N_HOLE_CARDS
    for e in range(min(12, g["deck"].shape[0])):
        d = g["deck"][e]
        csd = {"deck": {"deck_remaining": lut.get_2d_cards(d[2 * nh:])}, "board": None,
               "hand": [lut.get_2d_cards(d[p * nh:(p + 1) * nh]) for p in range(2)]}
        obs, rew, done, info = env.reset(deck_state_dict=csd)
        assert np.array_equal(obs, g["obs0"][e]) and not done and not np.any(rew)
        back = env.cards_state_dict()
        assert np.array_equal(back["hand"][0], csd["hand"][0]) and np.array_equal(back["deck"]["deck_remaining"], csd["deck"]["deck_remaining"])
        for t in range(int(g["n_steps"][e])):
            assert env.get_legal_actions() == list(np.nonzero(g["legal"][e, t])[0])
            if t == 1:  # snapshot, wander off, restore (PokerEnv.py:1161-1251)
                snap = env.state_dict()
                assert snap["seats"][0]["stack"] + snap["seats"][1]["stack"] + snap["main_pot"] + \
                    snap["seats"][0]["current_bet"] + snap["seats"][1]["current_bet"] == 2 * stack
                env.step(env.get_legal_actions()[-1])
                env.load_state_dict(snap)
                assert env.get_legal_actions() == list(np.nonzero(g["legal"][e, t])[0])
            obs, rew, done, info = env.step(int(g["action"][e, t]))
            assert np.array_equal(obs, g["obs"][e, t]) and np.array_equal(rew, g["rew"][e, t]) and done == bool(g["done"][e, t])


def test_run_cfrp_example_body_through_the_compat_namespace(tmp_path, golden_dir):
    """The body of examples/run_cfrp_example.py:17-37 - ChiefBase + CrayonWrapper + CFRPlus(DiscretizedNLLeduc, POT_ONLY) and
    150 x (iteration, update_from_log_buffer, export_all) - with `PokerRL` resolved to pokerrl_b200/compat: the exported log
    series equal the reference's own run of that script (tests/golden/cfr_CFRPlus_NLLeduc_POT.npz)."""
    code = '''
from PokerRL.cfr.CFRPlus import CFRPlus
from PokerRL.game import bet_sets
from PokerRL.game.games import DiscretizedNLLeduc
from PokerRL.rl.base_cls.workers.ChiefBase import ChiefBase
from PokerRL._.CrayonWrapper import CrayonWrapper
n_iterations = 150
name = "CFRplus_EXAMPLE"
chief = ChiefBase(t_prof=None)
crayon = CrayonWrapper(name=name, path_log_storage=%r, chief_handle=chief, runs_distributed=False, runs_cluster=False)
cfr = CFRPlus(name=name, game_cls=DiscretizedNLLeduc, delay=0, agent_bet_set=bet_sets.POT_ONLY, chief_handle=chief)
for iter_id in range(n_iterations):
    cfr.iteration()
    crayon.update_from_log_buffer()
    crayon.export_all(iter_nr=iter_id)
''' % str(tmp_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "pokerrl_b200", "compat"), ROOT]),
               PRL_AVG_F64="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    logs = json.load(open(os.path.join(str(tmp_path), "CFRplus_EXAMPLE", "149", "as_json", "logs.json")))
    g = np.load(os.path.join(golden_dir, "cfr_CFRPlus_NLLeduc_POT.npz"))
    curr = logs["CFRplus_EXAMPLE_Curr_S20000_total_CFRp_delay0"]["Evaluation/MBB_per_G"]
    avg = logs["CFRplus_EXAMPLE_Avg_total_S20000_CFRp_delay0"]["Evaluation/MBB_per_G"]
    curr = np.array([[int(k), v] for d in curr for k, v in d.items()])
    avg = np.array([[int(k), v] for d in avg for k, v in d.items()])
    assert np.array_equal(curr, g["curr_series"]) and np.array_equal(avg, g["avg_series"])
    assert "CFRplus_EXAMPLE_Curr_total_averaged_CFRp_delay0" in logs and "CFRplus_EXAMPLE_Avg_total_averaged_CFRp_delay0" in logs
