"""Evaluator base: experiment naming and result logging of `PokerRL/eval/_/EvaluatorMasterBase.py:9-172`
(same experiment / graph names; local calls instead of ray RPC).

Series per evaluation mode m and stack s of `t_prof`:   "<name> <m>_stack_<s>: <type> Total"
  with log_conf_interval additionally                   "... <type> Conf_lower95" / "... Conf_upper95"
several stack sizes: the average over them              "<name> <m>Multi_Stack: <type> Averaged Total"
  and its bounds                                        "<name> <m>: <type> Conf_lower95 / Conf_upper95"
all on the graph "Evaluation/<WIN_METRIC of the game>"."""
import numpy as np

_BOUNDS = ("lower95", "upper95")


class EvaluatorMasterBase:
    def __init__(self, t_prof, eval_env_bldr, chief_handle, eval_type, log_conf_interval=False):
        self._t_prof, self._eval_env_bldr, self._chief_handle = t_prof, eval_env_bldr, chief_handle
        self._chief_info = [None] * t_prof.n_seats
        self._is_multi_stack = len(t_prof.eval_stack_sizes) > 1
        self._log_conf_interval = bool(log_conf_interval)
        modes, new = t_prof.eval_modes_of_algo, chief_handle.create_experiment

        def per_stack(mode, suffix):
            return [new("%s %s_stack_%s: %s %s" % (t_prof.name, mode, stack[0], eval_type, suffix)) for stack in t_prof.eval_stack_sizes]

        self._exp_names_conf = None
        if self._log_conf_interval:  # [mode][stack] -> (lower, upper)
            self._exp_names_conf = {m: [list(pair) for pair in zip(*(per_stack(m, "Conf_" + b) for b in _BOUNDS))] for m in modes}
        self._exp_name_total = {m: per_stack(m, "Total") for m in modes}
        if self._is_multi_stack:
            self._exp_name_multi_stack = {m: new("%s %sMulti_Stack: %s Averaged Total" % (t_prof.name, m, eval_type)) for m in modes}
            if self._log_conf_interval:
                self._exp_names_multi_stack_conf = {m: [new("%s %s: %s Conf_%s" % (t_prof.name, m, eval_type, b)) for b in _BOUNDS]
                                                    for m in modes}

    is_multi_stack = property(lambda self: self._is_multi_stack)

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
        """(mean, half width of the normal-approximation 95 % interval of the mean)"""
        scores = np.asarray(scores)
        return float(scores.mean()), float(1.96 * scores.std() / np.sqrt(scores.shape[0]))

    def _emit(self, total_exp, bound_exps, iter_nr, score, lower_conf95, upper_conf95):
        g, add = self._graph(), self._chief_handle.add_scalar
        add(total_exp, g, iter_nr, score)
        if self._log_conf_interval:
            if lower_conf95 is None or upper_conf95 is None:
                raise ValueError("this evaluator logs confidence bounds: pass lower_conf95 and upper_conf95")
            for exp, v in zip(bound_exps, (lower_conf95, upper_conf95)):
                add(exp, g, iter_nr, v)

    def _log_results(self, agent_mode, stack_size_idx, iter_nr, score, upper_conf95=None, lower_conf95=None):
        bounds = self._exp_names_conf[agent_mode][stack_size_idx] if self._log_conf_interval else None
        self._emit(self._exp_name_total[agent_mode][stack_size_idx], bounds, iter_nr, score, lower_conf95, upper_conf95)

    def _log_multi_stack(self, agent_mode, iter_nr, score_total, upper_conf95=None, lower_conf95=None):
        bounds = self._exp_names_multi_stack_conf[agent_mode] if self._log_conf_interval else None
        self._emit(self._exp_name_multi_stack[agent_mode], bounds, iter_nr, score_total, lower_conf95, upper_conf95)
