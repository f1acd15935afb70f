"""Evaluator base: experiment naming and result logging of `PokerRL/eval/_/EvaluatorMasterBase.py:9-172`
(same experiment / graph names; local calls instead of ray RPC)."""


class EvaluatorMasterBase:
    def __init__(self, t_prof, eval_env_bldr, chief_handle, eval_type, log_conf_interval=False):
        self._t_prof = t_prof
        self._eval_env_bldr = eval_env_bldr
        self._chief_handle = chief_handle
        self._chief_info = [None for _ in range(t_prof.n_seats)]
        self._is_multi_stack = len(t_prof.eval_stack_sizes) > 1
        self._log_conf_interval = bool(log_conf_interval)
        bounds = ("lower95", "upper95")
        # "<name> <mode>_stack_<s>: <type> Conf_lower95 / Conf_upper95" (EvaluatorMasterBase.py:83-102)
        self._exp_names_conf = None if not log_conf_interval else {
            mode: [[chief_handle.create_experiment(t_prof.name + " " + mode + "_stack_" + str(stack[0]) + ": " + eval_type
                                                   + " Conf_" + b) for b in bounds] for stack in t_prof.eval_stack_sizes]
            for mode in t_prof.eval_modes_of_algo}
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
            if log_conf_interval:  # :40-54
                self._exp_names_multi_stack_conf = {
                    mode: [chief_handle.create_experiment(t_prof.name + " " + mode + ": " + eval_type + " Conf_" + b) for b in bounds]
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

    @staticmethod
    def _get_95confidence(scores):
        """(mean, half width of the normal 95 % interval) of a sample (EvaluatorMasterBase.py:123-128)"""
        import numpy as np
        scores = np.asarray(scores)
        return float(np.mean(scores)), float(1.96 * np.std(scores) / np.sqrt(scores.shape[0]))

    def _log_results(self, agent_mode, stack_size_idx, iter_nr, score, upper_conf95=None, lower_conf95=None):
        self._chief_handle.add_scalar(self._exp_name_total[agent_mode][stack_size_idx], self._graph(), iter_nr, score)
        if self._log_conf_interval:
            assert upper_conf95 is not None and lower_conf95 is not None
            lo, hi = self._exp_names_conf[agent_mode][stack_size_idx]
            self._chief_handle.add_scalar(lo, self._graph(), iter_nr, lower_conf95)
            self._chief_handle.add_scalar(hi, self._graph(), iter_nr, upper_conf95)

    def _log_multi_stack(self, agent_mode, iter_nr, score_total, upper_conf95=None, lower_conf95=None):
        self._chief_handle.add_scalar(self._exp_name_multi_stack[agent_mode], self._graph(), iter_nr, score_total)
        if self._log_conf_interval:
            assert upper_conf95 is not None and lower_conf95 is not None
            lo, hi = self._exp_names_multi_stack_conf[agent_mode]
            self._chief_handle.add_scalar(lo, self._graph(), iter_nr, lower_conf95)
            self._chief_handle.add_scalar(hi, self._graph(), iter_nr, upper_conf95)
