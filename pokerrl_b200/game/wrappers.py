"""Env builders with the attributes the CFR / BR path reads (`PokerRL/game/wrappers.py:18-68`,
`PokerRL/game/_/EnvWrapperBuilderBase.py:7-111`): `env_cls`, `env_args`, `rules`, `lut_holder`, `N_SEATS`,
`N_ACTIONS`.  The observation-history wrappers these builders create in the reference format NN inputs and are out
of scope."""
import copy


class EnvWrapperBuilderBase:
    def __init__(self, env_cls, env_args):
        self.env_cls = env_cls
        self.rules = env_cls.RULES
        self.env_args = env_args
        self._lut_holder = None
        self.N_SEATS = env_args.n_seats
        self.N_ACTIONS = env_args.N_ACTIONS

    @property
    def lut_holder(self):
        if self._lut_holder is None:
            self._lut_holder = self.env_cls.get_lut_holder()
        return self._lut_holder

    def args_for_stack(self, stack_size=None):
        args = copy.deepcopy(self.env_args)
        if stack_size is not None:
            assert isinstance(stack_size, list)
            args.starting_stack_sizes_list = copy.deepcopy(stack_size)
        return args


    def get_new_env(self, is_evaluating, stack_size=None):
        """EnvWrapperBuilderBase.get_new_env (:96-111): a single table on the device engine"""
        from pokerrl_b200.game.poker_env import PokerEnv
        return PokerEnv(self.env_cls, self.args_for_stack(stack_size), lut_holder=self.lut_holder, is_evaluating=is_evaluating)


class VanillaEnvBuilder(EnvWrapperBuilderBase):
    pass


class HistoryEnvBuilder(EnvWrapperBuilderBase):
    def __init__(self, env_cls, env_args, invert_history_order=False):
        super().__init__(env_cls=env_cls, env_args=env_args)
        self.invert_history_order = invert_history_order
