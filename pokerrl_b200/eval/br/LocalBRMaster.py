"""Exact best-response evaluator for any `EvalAgentBase` (`PokerRL/eval/br/LocalBRMaster.py:11-80`): one HBM-resident
`PublicTree` per evaluation stack size; `fill_with_agent_policy` -> reach pass -> value pass with BR on the GPU ->
`root.exploitability * EV_NORMALIZER` averaged over the two seats."""
import copy

from pokerrl_b200.eval._.EvaluatorMasterBase import EvaluatorMasterBase
from pokerrl_b200.game.PublicTree import PublicTree
from pokerrl_b200.rl.base_cls.TrainingProfileBase import get_env_builder


class LocalBRMaster(EvaluatorMasterBase):
    def __init__(self, t_prof, chief_handle, eval_agent_cls, device=None):
        super().__init__(t_prof=t_prof, eval_env_bldr=get_env_builder(t_prof=t_prof), chief_handle=chief_handle,
                         eval_type="BR")
        self._env_bldr = get_env_builder(t_prof=t_prof)
        assert self._env_bldr.N_SEATS == 2
        self._eval_agent = eval_agent_cls(t_prof=t_prof)
        self._game_trees = [PublicTree(env_bldr=self._env_bldr, stack_size=stack_size, stop_at_street=None,
                                       put_out_new_round_after_limit=True, is_debugging=t_prof.DEBUGGING, device=device)
                            for stack_size in t_prof.eval_stack_sizes]
        for gt in self._game_trees:
            gt.build_tree()
            print("Tree with stack size", gt.stack_size, "has", gt.n_nodes, "nodes out of which", gt.n_nonterm,
                  "are non-terminal.")

    @property
    def eval_agent(self):
        return self._eval_agent

    def evaluate(self, iter_nr):
        for mode in self._t_prof.eval_modes_of_algo:
            totals = []
            for stack_size_idx, stack_size in enumerate(self._t_prof.eval_stack_sizes):
                self._eval_agent.set_mode(mode)
                self._eval_agent.set_stack_size(stack_size=stack_size)
                if self._eval_agent.can_compute_mode():
                    e0, e1 = self._compute_br_heads_up(stack_size_idx=stack_size_idx, iter_nr=iter_nr)
                    self._log_results(iter_nr=iter_nr, agent_mode=mode, stack_size_idx=stack_size_idx,
                                      score=(e0 + e1) / 2)
                    totals.append((e0 + e1) / 2.0)
            if self._is_multi_stack and totals:
                self._log_multi_stack(agent_mode=mode, iter_nr=iter_nr, score_total=sum(totals) / float(len(totals)))

    def update_weights(self):
        self._eval_agent.update_weights(copy.deepcopy(self.pull_current_strat_from_chief()))

    def _compute_br_heads_up(self, stack_size_idx, iter_nr=None, do_export_tree=True):
        gt = self._game_trees[stack_size_idx]
        gt.fill_with_agent_policy(agent=self._eval_agent)
        gt.compute_ev()
        norm = self._env_bldr.env_cls.EV_NORMALIZER
        return float(gt.root.exploitability[0]) * norm, float(gt.root.exploitability[1]) * norm
