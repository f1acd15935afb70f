"""Linear CFR (`PokerRL/cfr/LinearCFR.py:9-76`): regrets and the reach-weighted strategy sum are weighted by the
iteration number.  Arithmetic: csrc/cfr_levels.cu."""
from pokerrl_b200.cfr._CFRBase import CFRBase as _CFRBase


class LinearCFR(_CFRBase):
    _SOLVER_ALGO = "LinearCFR"

    def __init__(self, name, chief_handle, game_cls, agent_bet_set, starting_stack_sizes=None, **engine_kw):
        super().__init__(name=name, chief_handle=chief_handle, game_cls=game_cls,
                         starting_stack_sizes=starting_stack_sizes, agent_bet_set=agent_bet_set,
                         algo_name="LinCFR", **engine_kw)
        self.reset()
