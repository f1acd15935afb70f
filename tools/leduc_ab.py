"""A/B of the one-card iteration schedules on one GPU (DiscretizedNLLeduc, bet_sets.B_5 by default).
usage: python tools/leduc_ab.py [workload=leduc_b5] [iterations=200]
Variants: level schedule in the persistent kernel at 512 threads per SM (default) and at 1024 threads / 64 registers
(PRL_PERSISTENT_THREADS=1024); subtree ("task") schedule at several thresholds.  For every variant: CFR+ iterations/s over
`iterations` iterations (CUDA events, after a warm-up call) and bit-equality of the regret table with the default."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from pokerrl_b200.solver import CFRSolver  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "leduc_b5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
t = time.time()
_, ft = bench.make_tree(workload, 20000)
print("tree", ft.n_nodes, "nodes %.1fs" % (time.time() - t), flush=True)

VARIANTS = [("levels, 512 threads/SM", {}, {}),
            ("levels, 1024 threads/SM", {"PRL_PERSISTENT_THREADS": "1024"}, {}),
            ("tasks, threshold 256", {}, dict(schedule="tasks", task_threshold=256)),
            ("tasks, threshold 1024", {}, dict(schedule="tasks", task_threshold=1024)),
            ("tasks, threshold 4096", {}, dict(schedule="tasks", task_threshold=4096))]
base = None
for name, env, kw in VARIANTS:
    os.environ.pop("PRL_PERSISTENT_THREADS", None)
    os.environ.update(env)
    s = CFRSolver(ft, "CFRPlus", **kw)
    s.iteration(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    while done < iters:
        n = min(20, iters - done)
        s.iteration(n)
        done += n
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    reg = s.bufs.regret.clone()
    line = "%-26s %8.1f iterations/s  (%.1f us per iteration)" % (name, iters / ms * 1e3, ms / iters * 1e3)
    if base is None:
        base = reg
    else:
        line += "  regrets %s" % ("bit-identical" if torch.equal(reg, base) else
                                  "DIFFER max %.3e" % float((reg - base).abs().max()))
    if kw.get("schedule") == "tasks":
        line += "  " + str(s._task_sched.stats())
    print(line, flush=True)
    del s, reg
    torch.cuda.empty_cache()
