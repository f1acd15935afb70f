"""Subtree ("task") schedule of the one-card sweeps - host side (plan for the one-card games).

The level-synchronous schedule needs one grid-wide step per tree level and sweep (54 per CFR iteration on the 14-level
Leduc B_5 tree).  Here the tree is cut by SUBTREE SIZE instead: every node whose subtree holds more than `threshold`
nodes belongs to the TRUNK; every other node whose parent is a trunk node roots a TASK = its complete subtree.  A task is
closed under the data flow of both sweeps (children -> parent bottom-up, parent -> children top-down), so one thread
block can run all its levels with block barriers; only the trunk (a few hundred nodes) needs the grid-wide order
tasks -> trunk (bottom-up) and trunk -> tasks (top-down).

Work lists (same conventions as FlatTree.work_order(): inside a level non-terminals first, sorted by kind / fan-out so
that a warp holds nodes of one kind):
    order        int32[n_nodes]      task-major: tasks by descending size, then local level, then the work-order key;
                                     the trunk's nodes follow, by tree level
    seg_start    int32[n_seg + 1]    a segment = one (task, local level); entries of segment s = order[seg_start[s] :
                                     seg_start[s + 1]]
    seg_nonterm  int32[n_seg]        non-terminal entries of the segment (they come first)
    task_ptr     int32[n_tasks + 1]  segments of task t = task_ptr[t] : task_ptr[t + 1]   (top level first)
    trunk_start  int64[n_levels + 1] trunk entries of tree level d = order[trunk_start[d] : trunk_start[d + 1]]
"""
import numpy as np

KIND_CHANCE = 2


def subtree_sizes(ft):
    size = np.ones(ft.n_nodes, np.int64)
    for d in range(ft.n_levels - 1, 0, -1):
        lo, hi = int(ft.level_start[d]), int(ft.level_start[d + 1])
        np.add.at(size, ft.parent[lo:hi], size[lo:hi])
    return size


def node_depths(ft):
    depth = np.zeros(ft.n_nodes, np.int32)
    for d in range(ft.n_levels):
        depth[int(ft.level_start[d]):int(ft.level_start[d + 1])] = d
    return depth


class TaskSchedule:
    def __init__(self, ft, threshold=1024):
        N = ft.n_nodes
        size, depth = subtree_sizes(ft), node_depths(ft)
        trunk = size > threshold
        par = ft.parent.astype(np.int64)
        is_root = ~trunk
        is_root[1:] &= trunk[par[1:]]  # a task root hangs below a trunk node (or is the tree's root: whole tree <= threshold)
        task_root = np.full(N, -1, np.int64)
        task_root[is_root] = np.nonzero(is_root)[0]
        for d in range(1, ft.n_levels):  # top-down: inherit the parent's task
            lo, hi = int(ft.level_start[d]), int(ft.level_start[d + 1])
            ids = np.arange(lo, hi)
            inherit = ~trunk[ids] & ~is_root[ids]
            task_root[ids[inherit]] = task_root[par[ids[inherit]]]
        assert np.all((task_root >= 0) == ~trunk)
        roots = np.nonzero(is_root)[0]
        roots = roots[np.argsort(-size[roots], kind="stable")]  # big tasks first: better packing of the block scheduler
        rank = np.full(N, -1, np.int64)
        rank[roots] = np.arange(roots.size)
        nonterm = ft.kind <= KIND_CHANCE  # FlatTree.work_order() counts them the same way (level_nonterm)
        # the work-order key of FlatTree.work_order(): non-terminals (kinds 0..2) before terminals, then chance depth / fan-out
        key = (ft.kind.astype(np.int64) << 40) + (ft.cdepth.astype(np.int64) << 32) + ft.n_children.astype(np.int64)
        in_task = np.nonzero(~trunk)[0]
        t_of = rank[task_root[in_task]]
        local = depth[in_task] - depth[task_root[in_task]]
        perm = np.lexsort((in_task, key[in_task], local, t_of))  # last key is the primary one
        task_nodes = in_task[perm]
        t_sorted, l_sorted = t_of[perm], local[perm]
        seg_id = t_sorted * (int(local.max()) + 1 if local.size else 1) + l_sorted
        new_seg = np.ones(task_nodes.size, bool)
        new_seg[1:] = seg_id[1:] != seg_id[:-1]
        seg_first = np.nonzero(new_seg)[0]
        self.seg_start = np.concatenate([seg_first, [task_nodes.size]]).astype(np.int32)
        seg_of = np.cumsum(new_seg) - 1
        self.seg_nonterm = np.bincount(seg_of, weights=nonterm[task_nodes], minlength=seg_first.size).astype(np.int32)
        seg_task = t_sorted[seg_first]
        self.n_tasks = int(roots.size)
        self.task_ptr = np.searchsorted(seg_task, np.arange(self.n_tasks + 1)).astype(np.int32)
        # trunk: by tree level, work-order key inside a level (every trunk node is a non-terminal)
        tr = np.nonzero(trunk)[0]
        tperm = np.lexsort((tr, key[tr], depth[tr]))
        trunk_nodes = tr[tperm]
        self.trunk_start = (task_nodes.size + np.searchsorted(depth[trunk_nodes], np.arange(ft.n_levels + 1))).astype(np.int64)
        self.order = np.concatenate([task_nodes, trunk_nodes]).astype(np.int32)
        self.n_task_nodes, self.n_trunk = int(task_nodes.size), int(trunk_nodes.size)
        self.task_roots, self.task_sizes = roots.astype(np.int64), size[roots]
        self.threshold = int(threshold)

    def stats(self):
        return {"threshold": self.threshold, "tasks": self.n_tasks, "trunk_nodes": self.n_trunk,
                "task_size_median": int(np.median(self.task_sizes)) if self.n_tasks else 0,
                "task_size_max": int(self.task_sizes.max()) if self.n_tasks else 0, "segments": int(self.seg_nonterm.size)}
