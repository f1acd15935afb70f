"""Logging sink of the CFR / BR path: in-memory experiment buffer with the reference's interface
(`PokerRL/rl/base_cls/workers/ChiefBase.py:7-109`: create_experiment / add_scalar / get_new_values).

The reference's Chief is also a ray actor base (WorkerBase.py:10-38); the tabular CFR path always runs it locally
(`ChiefBase(t_prof=None)`, examples/run_cfrp_example.py:21), so only the log buffer is provided here."""


class _LogBuffer:
    def __init__(self):
        self._experiments = {}
        self._new_values = {}

    def clear(self):
        self._experiments = {}

    def create_experiment(self, name):
        self._experiments.setdefault(name, {})
        return name

    def add_scalar(self, exp_name, graph_name, step, value):
        if exp_name not in self._experiments:
            raise AttributeError("Should create experiment before adding to it")
        self._experiments[exp_name].setdefault(graph_name, []).append([step, value])
        self._new_values.setdefault(exp_name, {}).setdefault(graph_name, []).append([step, value])

    def get_new_values(self):
        new_v, self._new_values = self._new_values, {}
        return new_v, list(self._experiments.keys())


class ChiefBase:
    def __init__(self, t_prof=None):
        self._t_prof = t_prof
        self._experiment_names = {}
        self._log_buf = _LogBuffer()

    def pull_current_eval_strategy(self, last_iteration_receiver_has):
        raise NotImplementedError

    def export_agent(self, step):
        raise NotImplementedError

    def create_experiment(self, name):
        return self._log_buf.create_experiment(name)

    def add_scalar(self, exp_name, graph_name, step, value):
        self._log_buf.add_scalar(exp_name=exp_name, graph_name=graph_name, step=step, value=value)

    def get_new_values(self):
        return self._log_buf.get_new_values()

    def get_experiments(self):
        """All logged series: {experiment: {graph: [[step, value], ...]}} (extension; the reference reads the
        private `_log_buf._experiments`)."""
        return self._log_buf._experiments
