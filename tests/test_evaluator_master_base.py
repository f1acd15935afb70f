"""Experiment names and logging of the evaluator base (`PokerRL/eval/_/EvaluatorMasterBase.py:83-172`) with a recording chief:
names per mode / stack, confidence bounds, multi-stack averages.  CPU only."""
from pokerrl_b200.eval._.EvaluatorMasterBase import EvaluatorMasterBase


class _TProf:
    name, n_seats = "run", 2
    eval_stack_sizes = [[100, 100], [200, 200]]
    eval_modes_of_algo = ("AVG", "CUR")


class _Chief:
    def __init__(self):
        self.names, self.log = [], []

    def create_experiment(self, name):
        self.names.append(name)
        return name

    def add_scalar(self, exp, graph, step, value):
        self.log.append((exp, graph, step, value))


class _Bldr:
    class env_cls:
        WIN_METRIC = "MBB_per_G"


def test_names_and_series_follow_the_reference():
    chief = _Chief()
    m = EvaluatorMasterBase(_TProf(), _Bldr(), chief, "LBR", log_conf_interval=True)
    assert m.is_multi_stack
    for want in ("run AVG_stack_100: LBR Total", "run CUR_stack_200: LBR Conf_lower95", "run AVG_stack_200: LBR Conf_upper95",
                 "run AVGMulti_Stack: LBR Averaged Total", "run CUR: LBR Conf_lower95", "run CUR: LBR Conf_upper95"):
        assert want in chief.names, want
    assert len(chief.names) == len(set(chief.names)) == 2 * 2 * 3 + 2 + 2 * 2
    m._log_results("AVG", 1, 5, 1.0, upper_conf95=2.0, lower_conf95=0.5)
    m._log_multi_stack("CUR", 5, 3.0, upper_conf95=4.0, lower_conf95=2.5)
    g = "Evaluation/MBB_per_G"
    assert chief.log == [("run AVG_stack_200: LBR Total", g, 5, 1.0), ("run AVG_stack_200: LBR Conf_lower95", g, 5, 0.5),
                         ("run AVG_stack_200: LBR Conf_upper95", g, 5, 2.0), ("run CURMulti_Stack: LBR Averaged Total", g, 5, 3.0),
                         ("run CUR: LBR Conf_lower95", g, 5, 2.5), ("run CUR: LBR Conf_upper95", g, 5, 4.0)]
    mean, half = m._get_95confidence([1.0, 2.0, 3.0, 4.0])
    assert abs(mean - 2.5) < 1e-12 and abs(half - 1.96 * (1.25 ** 0.5) / 2.0) < 1e-12


def test_without_confidence_bounds_only_totals_are_logged():
    chief = _Chief()
    m = EvaluatorMasterBase(_TProf(), _Bldr(), chief, "BR")
    assert all("Conf_" not in n for n in chief.names)
    m._log_results("AVG", 0, 1, 7.0)
    assert chief.log == [("run AVG_stack_100: BR Total", "Evaluation/MBB_per_G", 1, 7.0)]
