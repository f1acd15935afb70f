"""Evaluator base: experiment naming and result logging of `PokerRL/eval/_/EvaluatorMasterBase.py:9-172`
(same experiment / graph names; local calls instead of ray RPC)."""


class EvaluatorMasterBase:
    def __init__(self, t_prof, eval_env_bldr, chief_handle, eval_type, log_conf_interval=False):
        self._t_prof = t_prof
        self._eval_env_bldr = eval_env_bldr
        self._chief_handle = chief_handle
        self._chief_info = [None for _ in range(t_prof.n_seats)]
        self._is_multi_stack = len(t_prof.eval_stack_sizes) > 1
        self._exp_name_total = {
            mode: [chief_handle.create_experiment(
                t_prof.name + " " + mode + "_stack_" + str(stack[0]) + ": " + eval_type + " Total")
                for stack in t_prof.eval_stack_sizes]
            for mode in t_prof.eval_modes_of_algo}
        if self._is_multi_stack:
            self._exp_name_multi_stack = {
                mode: chief_handle.create_experiment(
                    t_prof.name + " " + mode + "Multi_Stack" + ": " + eval_type + " Averaged Total")
                for mode in t_prof.eval_modes_of_algo}

    @property
    def is_multi_stack(self):
        return self._is_multi_stack

    def evaluate(self, iter_nr):
        raise NotImplementedError

    def update_weights(self):
        raise NotImplementedError

    def pull_current_strat_from_chief(self):
        w, self._chief_info = self._chief_handle.pull_current_eval_strategy(self._chief_info)
        return w

    def _graph(self):
        return "Evaluation/" + self._eval_env_bldr.env_cls.WIN_METRIC

    def _log_results(self, agent_mode, stack_size_idx, iter_nr, score, upper_conf95=None, lower_conf95=None):
        self._chief_handle.add_scalar(self._exp_name_total[agent_mode][stack_size_idx], self._graph(), iter_nr, score)

    def _log_multi_stack(self, agent_mode, iter_nr, score_total, upper_conf95=None, lower_conf95=None):
        self._chief_handle.add_scalar(self._exp_name_multi_stack[agent_mode], self._graph(), iter_nr, score_total)
