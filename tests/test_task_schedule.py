"""Host checks of the subtree ("task") schedule (pokerrl_b200/task_schedule.py): it must be a re-ordering of the same
per-node work that respects the data flow of both sweeps - children before parents bottom-up, parents before children
top-down - with only block-local ordering inside a task and the grid-wide order tasks -> trunk / trunk -> tasks."""
import numpy as np
import pytest

from common import make_flat_tree
from pokerrl_b200.task_schedule import TaskSchedule, subtree_sizes


@pytest.mark.parametrize("game,threshold", [("StandardLeduc", 40), ("NLLeduc_POT", 64), ("NLLeduc_B3", 512),
                                            ("StandardLeduc", 10 ** 6), ("NLLeduc_POT", 1)])
def test_schedule_is_a_valid_reordering(game, threshold):
    ft = make_flat_tree(game)
    ts = TaskSchedule(ft, threshold)
    N = ft.n_nodes
    assert sorted(ts.order.tolist()) == list(range(N))
    assert ts.n_task_nodes + ts.n_trunk == N and ts.seg_start[-1] == ts.n_task_nodes
    size = subtree_sizes(ft)
    in_trunk = np.zeros(N, bool)
    in_trunk[ts.order[ts.n_task_nodes:]] = True
    assert np.array_equal(in_trunk, size > threshold)
    if threshold >= N:
        assert ts.n_tasks == 1 and ts.n_trunk == 0
    # ---- bottom-up: tasks (each from its deepest segment to its root), then the trunk from the deepest tree level
    done = np.zeros(N, bool)
    task_of = np.full(N, -1)

    def run_value(nodes):
        for n in nodes:
            fc, A = ft.first_child[n], ft.n_children[n]
            if fc >= 0:
                assert done[fc:fc + A].all(), ("child not ready", n)
        done[nodes] = True  # a whole segment runs between two barriers

    for t in range(ts.n_tasks):
        s0, s1 = ts.task_ptr[t], ts.task_ptr[t + 1]
        assert s1 > s0
        root_seg = ts.order[ts.seg_start[s0]:ts.seg_start[s0 + 1]]
        assert root_seg.size == 1 and root_seg[0] == ts.task_roots[t]
        for s in range(s0, s1):
            task_of[ts.order[ts.seg_start[s]:ts.seg_start[s + 1]]] = t
        for s in range(s1 - 1, s0 - 1, -1):
            seg = ts.order[ts.seg_start[s]:ts.seg_start[s + 1]]
            k = ft.kind[seg]
            assert (k[:ts.seg_nonterm[s]] <= 2).all() and (k[ts.seg_nonterm[s]:] >= 3).all()
            assert (np.diff(k.astype(int)) >= 0).all()  # warps see one kind at a time
            run_value(seg)
        assert int((task_of == t).sum()) == size[ts.task_roots[t]]  # the task is the root's complete subtree
    for d in range(ft.n_levels - 1, -1, -1):
        run_value(ts.order[ts.trunk_start[d]:ts.trunk_start[d + 1]])
    assert done.all()
    # ---- top-down: the trunk level by level (parents write their children's rows), then the tasks
    have_row = np.zeros(N, bool)
    have_row[0] = True  # reach_group writes the root's own row when it processes the root

    def run_reach(parents):
        for n in parents:
            assert have_row[n], ("parent row missing", n)
        for n in parents:
            fc, A = ft.first_child[n], ft.n_children[n]
            if fc >= 0:
                have_row[fc:fc + A] = True

    for d in range(ft.n_levels):
        run_reach(ts.order[ts.trunk_start[d]:ts.trunk_start[d + 1]])
    for t in range(ts.n_tasks):
        for s in range(ts.task_ptr[t], ts.task_ptr[t + 1]):
            run_reach(ts.order[ts.seg_start[s]:ts.seg_start[s] + ts.seg_nonterm[s]])
    assert have_row.all()
    # parents of task roots are trunk nodes; all other task nodes have their parent in the same task
    for t in range(ts.n_tasks):
        r = ts.task_roots[t]
        assert r == 0 or in_trunk[ft.parent[r]]
    inner = (task_of >= 0) & ~np.isin(np.arange(N), ts.task_roots)
    assert np.array_equal(task_of[inner], task_of[ft.parent[inner]])


def test_b5_statistics_quoted_in_design_md():
    import bench
    _, ft = bench.make_tree("leduc_b5", 20000)
    st = TaskSchedule(ft, 1024).stats()
    assert (st["trunk_nodes"], st["tasks"], st["task_size_max"]) == (551, 3037, 1011)


@pytest.mark.parametrize("algo", ["CFRPlus", "LinearCFR", "VanillaCFR"])
@pytest.mark.parametrize("game,threshold", [("NLLeduc_POT", 64), ("NLLeduc_POT", 10 ** 6), ("StandardLeduc", 20),
                                            ("NLLeduc_POT", 0)])
def test_task_order_reproduces_level_order_bit_for_bit(algo, game, threshold):
    """The orchestration of prl_cfr_iterations_tasks restated on the host (oracle/cfr_oracle.c:
    orc_cfr_iterations_tasks, same launch sequence and pending-sweep bookkeeping as run_task_iterations in
    pokerrl_b200/csrc/cfr_levels.cu) against the level order, which the reference's fixtures pin."""
    import cfr_c
    ft = make_flat_tree(game)
    a = cfr_c.OracleCSolver(ft, algo, avg_f64=(algo == "CFRPlus"), n_threads=1)
    b = cfr_c.OracleCSolver(ft, algo, avg_f64=(algo == "CFRPlus"), n_threads=1)
    b.set_task_schedule(threshold)
    for n in (1, 1, 3):
        a.iteration(n)
        b.iteration(n)
        for k in ("regret", "strat", "avg", "reach"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (k, a.iter_counter)
        assert a.exploitability_current() == b.exploitability_current()
        assert a.exploitability_average() == b.exploitability_average()
