"""CFR+ (`PokerRL/cfr/CFRPlus.py:9-87`): regrets floored at 0, regret matching, linear (not reach-weighted)
averaging with `delay`.  Arithmetic: csrc/cfr_levels.cu."""
from pokerrl_b200.cfr._CFRBase import CFRBase as _CFRBase


class CFRPlus(_CFRBase):
    _SOLVER_ALGO = "CFRPlus"

    def __init__(self, name, chief_handle, game_cls, agent_bet_set, starting_stack_sizes=None, delay=0, **engine_kw):
        super().__init__(name=name, chief_handle=chief_handle, game_cls=game_cls,
                         starting_stack_sizes=starting_stack_sizes, agent_bet_set=agent_bet_set,
                         algo_name="CFRp_delay" + str(delay), delay=delay, **engine_kw)
        self.delay = delay
        self.reset()

    def _evaluate_avg_strats(self):
        if self._iter_counter > self.delay:  # CFRPlus.py:33-35
            return super()._evaluate_avg_strats()
