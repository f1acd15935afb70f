"""Replay of the REFERENCE's LBR episodes (tests/golden/lbr_runs.npz, produced by running PokerRL/eval/lbr/LocalLBRWorker.py
against a fixed hand-dependent policy: oracle/gen_golden_lbr_run.py) through this package's LBR worker on the GPU: same deals,
same agent random numbers -> LBR must take the same decisions and win the same chips in every hand; its utility vectors
(roll-out equities from csrc/lbr_rollout.cu in the reference's board-counter mode) agree to float32 round-off of the
reference's own float32 sums over up to 990 board completions (3.3e-5 of the largest utility)."""
import os

import numpy as np
import pytest

from lbr_common import ReplayTableAgent

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lbr_runs.npz"))


def _worker(game_name, bet_set):
    from pokerrl_b200.eval.lbr.LBRArgs import LBRArgs
    from pokerrl_b200.eval.lbr.LocalLBRWorker import LocalLBRWorker
    from pokerrl_b200.game import games
    from pokerrl_b200.game.Poker import Poker
    from pokerrl_b200.rl.base_cls.TrainingProfileBase import TrainingProfileBase
    game = getattr(games, game_name)
    t_prof = TrainingProfileBase("lbr", game, bet_set, eval_stack_sizes=[[20000, 20000]], eval_modes_of_algo=("table",))
    t_prof.env_builder_cls_str = "VanillaEnvBuilder"
    t_prof.module_args["lbr"] = LBRArgs(lbr_bet_set=bet_set, n_lbr_hands_per_seat=1, lbr_check_to_round=Poker.FLOP)
    return LocalLBRWorker(t_prof=t_prof, chief_handle=None, eval_agent_cls=ReplayTableAgent, reference_board_counter_quirk=True)


@pytest.mark.parametrize("game_name", ["Flop5Holdem", "DiscretizedNLHoldem"])
def test_reference_lbr_episodes_replay(game_name):
    from pokerrl_b200.game import bet_sets
    bet_set = bet_sets.B_2 if game_name == "DiscretizedNLHoldem" else [1.0]
    w = _worker(game_name, bet_set)
    decks, draws = GOLD[game_name + "_decks"], GOLD[game_name + "_draws"]
    utils, seats, wins = GOLD[game_name + "_utils"], GOLD[game_name + "_agent_seat"], GOLD[game_name + "_winnings"]
    w.agent.set_mode("table")
    w.agent.set_stack_size([20000, 20000])
    w._env = w._env_bldr.get_new_env(is_evaluating=True, stack_size=[20000, 20000])
    lut = w._env_bldr.lut_holder
    n_dec, worst, mism = 0, 0.0, []
    for i in range(len(decks)):
        d = decks[i]
        csd = {"hand": [lut.get_2d_cards(d[0:2]), lut.get_2d_cards(d[2:4])], "deck": {"deck_remaining": lut.get_2d_cards(d[4:])}}
        w.agent.draws = [float(u) for u in draws[i] if u >= 0]
        got = float(w.play_hand(int(seats[i]), csd))
        assert not w.agent.draws, "the replay consumed fewer random numbers than the reference (hand %d)" % i
        ref_u = [u[~np.isnan(u)] for u in utils[i] if not np.all(np.isnan(u))]
        assert len(ref_u) == len(w.last_utilities), (i, len(ref_u), len(w.last_utilities))
        for a, b in zip(w.last_utilities, ref_u):
            n_dec += 1
            assert a.shape == b.shape and np.array_equal(a == -1, b == -1), (i, a, b)
            worst = max(worst, float(np.abs(a - b).max() / max(1.0, np.abs(b).max())))
            assert int(np.argmax(a)) == int(np.argmax(b)), (i, a, b)
        if abs(got - wins[i]) > 1e-3 * max(1.0, abs(wins[i])):
            mism.append((i, got, float(wins[i])))
    print("%s: %d hands, %d LBR decisions replayed; utilities within %.1e of the reference's; winnings mismatches: %s"
          % (game_name, len(decks), n_dec, worst, mism))
    assert not mism and worst <= 1e-4 and n_dec > 50  # measured: Flop5Holdem (complete boards) ~1e-7, NL flop / turn roll-outs 3.3e-5


def test_lbr_master_logs_mean_and_confidence_interval():
    """LocalLBRMaster with one local worker: n hands per seat with fresh deals, mean and the 95 % bounds logged under the
    reference's experiment names (EvaluatorMasterBase.py:83-128); the policy agent samples with numpy's global RNG"""
    import numpy as np
    from lbr_common import policy_table
    from pokerrl_b200.eval.lbr.LBRArgs import LBRArgs
    from pokerrl_b200.eval.lbr.LocalLBRMaster import LocalLBRMaster
    from pokerrl_b200.eval.lbr.LocalLBRWorker import LocalLBRWorker
    from pokerrl_b200.game import games
    from pokerrl_b200.game.Poker import Poker
    from pokerrl_b200.rl.base_cls.EvalAgentBase import EvalAgentBase
    from pokerrl_b200.rl.base_cls.TrainingProfileBase import TrainingProfileBase
    from pokerrl_b200.rl.base_cls.workers.ChiefBase import ChiefBase

    class Agent(EvalAgentBase):
        ALL_MODES = ["table"]

        def can_compute_mode(self):
            return True

        def update_weights(self, w):
            pass

        def get_a_probs_for_each_hand(self):
            env = self.internal_env
            return policy_table(self.env_bldr.rules.RANGE_SIZE, self.env_bldr.N_ACTIONS, env.get_legal_actions(), env.current_round)

    t_prof = TrainingProfileBase("m", games.Flop5Holdem, [1.0], eval_stack_sizes=[[20000, 20000]], eval_modes_of_algo=("table",))
    t_prof.module_args["lbr"] = LBRArgs(n_lbr_hands_per_seat=30, lbr_check_to_round=Poker.FLOP)
    chief = ChiefBase(t_prof=None)
    chief.pull_current_eval_strategy = lambda info: (None, info)
    master = LocalLBRMaster(t_prof=t_prof, chief_handle=chief)
    master.set_worker_handles(LocalLBRWorker(t_prof=t_prof, chief_handle=chief, eval_agent_cls=Agent))
    np.random.seed(3)
    master.update_weights()
    master.evaluate(iter_nr=7)
    exps = chief.get_experiments()
    g = "Evaluation/" + games.Flop5Holdem.WIN_METRIC
    total = exps["m table_stack_20000: LBR Total"][g]
    lo, hi = exps["m table_stack_20000: LBR Conf_lower95"][g], exps["m table_stack_20000: LBR Conf_upper95"][g]
    assert total[0][0] == 7 and lo[0][1] <= total[0][1] <= hi[0][1] and hi[0][1] > lo[0][1]
    print("LBR master: %.1f [%0.1f, %0.1f] %s over 60 hands" % (total[0][1], lo[0][1], hi[0][1], games.Flop5Holdem.WIN_METRIC))
