"""LBR evaluator master (`PokerRL/eval/lbr/LocalLBRMaster.py:14-90`): for every evaluation mode and stack size it has its
worker(s) play `n_lbr_hands` hands per seat against the agent and logs LBR's mean winnings with the 95 % confidence interval
under the reference's experiment names ("<name> <mode>_stack_<s>: LBR Total / Conf_lower95 / Conf_upper95").  Workers are local
objects (eval/lbr/LocalLBRWorker.py: the hands run on the device engine); the reference's ray fan-out is out of scope."""
import numpy as np

from pokerrl_b200.eval._.EvaluatorMasterBase import EvaluatorMasterBase
from pokerrl_b200.rl.base_cls.TrainingProfileBase import get_env_builder


class LocalLBRMaster(EvaluatorMasterBase):
    def __init__(self, t_prof, chief_handle):
        assert t_prof.n_seats == 2
        super().__init__(t_prof=t_prof, eval_env_bldr=get_env_builder(t_prof), chief_handle=chief_handle, eval_type="LBR",
                         log_conf_interval=True)
        self.lbr_args = t_prof.module_args["lbr"]
        self.weights_for_eval_agent = None
        self.alive_worker_handles = None

    def set_worker_handles(self, *worker_handles):
        self.alive_worker_handles = list(worker_handles)

    def update_weights(self):
        self.weights_for_eval_agent = self.pull_current_strat_from_chief()

    def evaluate(self, iter_nr):
        workers = self.alive_worker_handles
        for w in workers:
            w.update_weights(self.weights_for_eval_agent)
        hands_per_worker = int(self.lbr_args.n_lbr_hands / max(1, len(workers)))
        for mode in self._t_prof.eval_modes_of_algo:
            means, halves = [], []
            for k, stack in enumerate(self._t_prof.eval_stack_sizes):
                scores = [w.run(seat, hands_per_worker, mode, stack) for seat in range(self._t_prof.n_seats) for w in workers]
                scores = [x for x in scores if x is not None]
                if not scores:
                    continue
                mean, half = self._get_95confidence(np.concatenate(scores, axis=0))
                self._log_results(agent_mode=mode, stack_size_idx=k, iter_nr=iter_nr, score=mean, upper_conf95=mean + half,
                                  lower_conf95=mean - half)
                means.append(mean)
                halves.append(half)
            if self._is_multi_stack and means:
                m, h = sum(means) / len(means), sum(halves) / len(halves)
                self._log_multi_stack(agent_mode=mode, iter_nr=iter_nr, score_total=m, upper_conf95=m + h, lower_conf95=m - h)
