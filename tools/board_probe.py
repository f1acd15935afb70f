"""Timing probe of the board engine on one GPU: python tools/board_probe.py [n_boards] [iterations] [grid]
CUDA-event times of the update sweeps (per seat), the trunk part and the evaluation passes."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pokerrl_b200.board_engine import BoardCFRSolver  # noqa: E402
from pokerrl_b200.game import games  # noqa: E402
from pokerrl_b200.game.holdem_boards import BoardSpec  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = games.Flop5Holdem
args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000], bet_sizes_list_as_frac_of_pot=[1.0])
t0 = time.perf_counter()
spec = BoardSpec.full_game(g.RULES)
if nb < spec.boards.shape[0]:
    spec = BoardSpec(spec.boards[:nb], spec.board_prob[:nb], spec.board_mult[:nb], spec.sym_perm, "first %d" % nb)
t_spec = time.perf_counter() - t0
t0 = time.perf_counter()
s = BoardCFRSolver(g, args, spec, grid=grid)
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
s.iteration(3)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
s.iteration(iters)
ev[1].record()
torch.cuda.synchronize()
it_ms = ev[0].elapsed_time(ev[1]) / iters
# sweep kernel alone
sw = []
for rep in range(6):
    for p in (0, 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s._sweep_begin(s.bufs, p, False, 0, 0)
        e1.record()
        torch.cuda.synchronize()
        sw.append(e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
a = s.exploitability_current()
b = s.exploitability_average()
e1.record()
torch.cuda.synchronize()
L = s.L
bytes_sweep = s.n_boards * (35 * L["ldb"] * 4 + L["blob"])
print(json.dumps({"boards": s.n_boards, "spec_s": t_spec, "build_s": t_build, "ms_per_iteration": it_ms,
                  "iterations_per_s": 1e3 / it_ms, "sweep_ms": sorted(sw)[len(sw) // 2], "sweep_ms_all": sw,
                  "sweep_GBps": bytes_sweep / (sorted(sw)[len(sw) // 2] * 1e-3) / 1e9, "bytes_per_sweep": bytes_sweep,
                  "eval_both_ms": e0.elapsed_time(e1), "expl_cur": a, "expl_avg": b,
                  "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30, "grid": s.g.grid}))
