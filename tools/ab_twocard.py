"""A/B of the two-card sweep kernel variants on one GPU (Flop5Holdem, first N board classes).
usage: python tools/ab_twocard.py [n_boards=20000] [iterations=6] [records-only]
Variants are selected by the switches the library / solver read: PRL_TERMINAL_V (terminal kernel generation 2 / 3 / 4), PRL_NO_HAND_REC
(no packed per-hand record), PRL_NO_NODE_REC (old tiled row kernels with pointer chains).  For every variant: mean value-
sweep and reach-sweep time per seat over the timed iterations, and max |regret| / exploitability differences against
the first (baseline) variant after the same number of iterations."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pokerrl_b200 import _native as nat
from pokerrl_b200.game import games
from pokerrl_b200.game.flat_tree import FlatTree
from pokerrl_b200.game.holdem_boards import BoardSpec
from pokerrl_b200.solver import CFRSolver, _stream

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
g = games.Flop5Holdem
args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000] * 2, bet_sizes_list_as_frac_of_pot=[1.0])
spec = BoardSpec.full_game(g.RULES)
if nb < spec.boards.shape[0]:
    spec = BoardSpec(spec.boards[:nb], spec.board_prob[:nb], spec.board_mult[:nb], spec.sym_perm, "first %d classes" % nb)
t = time.time()
ft = FlatTree(g, args, board_spec=spec)
print("tree", ft.n_nodes, "nodes %.1fs" % (time.time() - t), flush=True)

VARIANTS = [
    ("no records (fallbacks)", dict(PRL_TERMINAL_V="2", PRL_NO_HAND_REC="1", PRL_NO_NODE_REC="1")),
    ("v2 terminal + records", dict(PRL_TERMINAL_V="2", PRL_NO_HAND_REC="0", PRL_NO_NODE_REC="0")),
    ("v3 terminal + records", dict(PRL_TERMINAL_V="3", PRL_NO_HAND_REC="0", PRL_NO_NODE_REC="0")),
    ("v4 fold / showdown apart", dict(PRL_TERMINAL_V="4", PRL_NO_HAND_REC="0", PRL_NO_NODE_REC="0")),
]
if len(sys.argv) > 3 and sys.argv[3] == "records-only":
    VARIANTS = VARIANTS[1:]
if len(sys.argv) > 3 and sys.argv[3] == "newest":
    VARIANTS = VARIANTS[2:]
base = None
for name, env in VARIANTS:
    os.environ.update(env)
    s = CFRSolver(ft, "CFRPlus")
    tree, buf = C.byref(s.dtree.desc), C.byref(s.bufs.desc)
    v_ms, r_ms = [], []
    for it in range(iters):
        for p in (0, 1):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            nat.call("prl_cfr_sweep", tree, buf, s.algo, p, s.iter_counter, s.delay, 0, nat.modes(*s.modes), 1, _stream())
            e[1].record()
            nat.call("prl_cfr_sweep", tree, buf, s.algo, p, s.iter_counter, s.delay, 0, nat.modes(*s.modes), 2, _stream())
            e[2].record()
            torch.cuda.synchronize()
            s.modes[p] = nat.STRAT_F32
            if it >= 2:
                v_ms.append(e[0].elapsed_time(e[1]))
                r_ms.append(e[1].elapsed_time(e[2]))
        s.iter_counter += 1
    expl = s.exploitability_current()
    reg = s.bufs.regret.clone()
    line = "%-24s value sweep %.3f ms  reach sweep %.3f ms  iteration %.3f ms  expl %.6f" % (
        name, sum(v_ms) / len(v_ms), sum(r_ms) / len(r_ms), 2 * (sum(v_ms) / len(v_ms) + sum(r_ms) / len(r_ms)), expl)
    if base is None:
        base = (reg, expl)
    else:
        line += "  max|dregret| %.3e (max|regret| %.3e)  dexpl %.3e" % (
            float((reg - base[0]).abs().max()), float(base[0].abs().max()), expl - base[1])
    print(line, flush=True)
    del s, reg
    torch.cuda.empty_cache()
