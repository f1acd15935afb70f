"""Vanilla CFR (`PokerRL/cfr/VanillaCFR.py:9-77`): unweighted regrets, reach-weighted strategy sum.
Arithmetic: csrc/cfr_levels.cu."""
from pokerrl_b200.cfr._CFRBase import CFRBase as _CFRBase


class VanillaCFR(_CFRBase):
    _SOLVER_ALGO = "VanillaCFR"

    def __init__(self, name, chief_handle, game_cls, agent_bet_set, starting_stack_sizes=None, **engine_kw):
        super().__init__(name=name, chief_handle=chief_handle, game_cls=game_cls,
                         starting_stack_sizes=starting_stack_sizes, agent_bet_set=agent_bet_set,
                         algo_name="CFR", **engine_kw)
        self.reset()
