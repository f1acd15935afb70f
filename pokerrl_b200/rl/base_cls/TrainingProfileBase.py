"""Minimal training profile: the fields the evaluators of the CFR / BR path read
(`PokerRL/rl/base_cls/TrainingProfileBase.py:16-139`; read sites `eval/br/LocalBRMaster.py:19-33`,
`eval/_/EvaluatorMasterBase.py:21-42`, `rl/rl_util.py:82-84`).  The deep-RL / ray / checkpoint-path plumbing of the
reference's profile is out of scope."""
import copy

from pokerrl_b200.game.games import get_env_cls_from_str
from pokerrl_b200.game.wrappers import HistoryEnvBuilder


class TrainingProfileBase:
    def __init__(self, name, game_cls, agent_bet_set, eval_stack_sizes=None, eval_modes_of_algo=("AVG",),
                 n_seats=2, DEBUGGING=False, device_inference="cuda:0"):
        self.name = name
        self.n_seats = n_seats
        self.game_cls_str = game_cls.__name__
        self.env_builder_cls_str = "HistoryEnvBuilder"
        self.eval_modes_of_algo = tuple(eval_modes_of_algo)
        self.eval_stack_sizes = copy.deepcopy(eval_stack_sizes) or [[game_cls.DEFAULT_STACK_SIZE] * n_seats]
        self.DEBUGGING, self.DISTRIBUTED, self.CLUSTER = DEBUGGING, False, False
        self.device_inference = device_inference
        self.module_args = {"env": game_cls.ARGS_CLS(n_seats=n_seats,
                                                     starting_stack_sizes_list=list(self.eval_stack_sizes[0]),
                                                     bet_sizes_list_as_frac_of_pot=list(agent_bet_set))}


def get_env_builder(t_prof):
    """rl_util.get_env_builder (rl_util.py:82-84)"""
    return HistoryEnvBuilder(env_cls=get_env_cls_from_str(t_prof.game_cls_str), env_args=t_prof.module_args["env"])
