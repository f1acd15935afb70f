"""GPU check of the subtree ("task") schedule of the one-card sweeps (prl_cfr_iterations_tasks): the same per-node
arithmetic in a different order of independent nodes must reproduce the level schedule BIT-FOR-BIT.

The schedule is experimental and not the default: it was written after the round-1 GPU budget was spent,
so this module only runs when PRL_TEST_TASKS=1 is set (first thing to do with a GPU: run it, then `python bench.py
--workload leduc_b5 --schedule tasks`)."""
import os

import pytest
import torch

from common import make_flat_tree

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PRL_TEST_TASKS") != "1", reason="experimental schedule: set PRL_TEST_TASKS=1")]


@pytest.mark.parametrize("algo", ["CFRPlus", "LinearCFR", "VanillaCFR"])
@pytest.mark.parametrize("name,threshold", [("NLLeduc_POT", 64), ("NLLeduc_POT", 10 ** 6), ("StandardLeduc", 20),
                                            ("NLLeduc_B3", 1024), ("NLLeduc_POT", 0)])
def test_task_schedule_equals_level_schedule(algo, name, threshold):
    from pokerrl_b200.solver import CFRSolver
    ft = make_flat_tree(name)
    a = CFRSolver(ft, algo, avg_f64=(algo == "CFRPlus"))
    b = CFRSolver(ft, algo, avg_f64=(algo == "CFRPlus"), schedule="tasks", task_threshold=threshold)
    for n in (1, 1, 3):  # single iterations and a multi-iteration call (pending top-down sweep across calls)
        a.iteration(n)
        b.iteration(n)
        for k in ("regret", "strat", "avg", "reach"):
            assert torch.equal(getattr(a.bufs, k), getattr(b.bufs, k)), (k, a.iter_counter)
        assert a.exploitability_current() == b.exploitability_current()
        assert a.exploitability_average() == b.exploitability_average()


def test_wide_persistent_kernel_equals_default(monkeypatch):
    """PRL_PERSISTENT_THREADS=1024: the persistent iteration kernel at 1024 threads / 64 registers per SM"""
    from pokerrl_b200.solver import CFRSolver
    ft = make_flat_tree("NLLeduc_B3")
    a = CFRSolver(ft, "CFRPlus", avg_f64=True)
    a.iteration(4)
    monkeypatch.setenv("PRL_PERSISTENT_THREADS", "1024")
    b = CFRSolver(ft, "CFRPlus", avg_f64=True)
    b.iteration(4)
    for k in ("regret", "strat", "avg", "reach"):
        assert torch.equal(getattr(a.bufs, k), getattr(b.bufs, k)), k
