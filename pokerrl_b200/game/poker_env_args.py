"""Environment argument objects with the constructor surface of `PokerRL/game/poker_env_args.py:4-131`.

`CFRBase` always forwards `bet_sizes_list_as_frac_of_pot` (`_CFRBase.py:50-56`); the limit / no-limit variants
swallow it through **kwargs exactly like the reference does."""


class _PokerEnvArgs:
    N_ACTIONS = 3

    def __init__(self, n_seats, starting_stack_sizes_list=None, stack_randomization_range=(0, 0),
                 scale_rewards=True, use_simplified_headsup_obs=True, return_pre_transition_state_in_info=False,
                 *args, **kwargs):
        self.n_seats = n_seats
        self.starting_stack_sizes_list = (list(starting_stack_sizes_list) if starting_stack_sizes_list is not None
                                          else [None] * n_seats)
        self.stack_randomization_range = stack_randomization_range
        self.scale_rewards = scale_rewards
        self.use_simplified_headsup_obs = use_simplified_headsup_obs
        self.RETURN_PRE_TRANSITION_STATE_IN_INFO = return_pre_transition_state_in_info


class NoLimitPokerEnvArgs(_PokerEnvArgs):
    pass


class LimitPokerEnvArgs(_PokerEnvArgs):
    pass


class DiscretizedPokerEnvArgs(_PokerEnvArgs):
    def __init__(self, n_seats, bet_sizes_list_as_frac_of_pot, starting_stack_sizes_list=None,
                 stack_randomization_range=(0, 0), uniform_action_interpolation=False,
                 use_simplified_headsup_obs=True, scale_rewards=True, return_pre_transition_state_in_info=False,
                 *args, **kwargs):
        super().__init__(n_seats, starting_stack_sizes_list, stack_randomization_range, scale_rewards,
                         use_simplified_headsup_obs, return_pre_transition_state_in_info)
        self.bet_sizes_list_as_frac_of_pot = list(bet_sizes_list_as_frac_of_pot)
        self.uniform_action_interpolation = uniform_action_interpolation
        self.N_ACTIONS = len(self.bet_sizes_list_as_frac_of_pot) + 2  # + FOLD and CHECK/CALL
