"""debug: one CFR+ iteration of the board engine vs the float64 C oracle, errors per local node"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cfr2_c  # noqa: E402
from pokerrl_b200.board_engine import BoardCFRSolver  # noqa: E402
from pokerrl_b200.game import games  # noqa: E402
from twocard_common import fhp_tree, oracle_tree, random_board_spec  # noqa: E402

spec = random_board_spec(48, 21)
ft = fhp_tree(spec)
orc = cfr2_c.Oracle2CSolver(ft, oracle_tree(ft).board_ranks, "CFRPlus", n_threads=8, lean=True)
g = games.Flop5Holdem
args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000], bet_sizes_list_as_frac_of_pot=[1.0])
s = BoardCFRSolver(g, args, spec)
st = ft.board_subtree()
for it in range(2):
    s.iteration(1)
    orc.iteration(1)
    reg, avg = [t.cpu().numpy()[:, :ft.R].astype(np.float64) for t in s.natural_tables(ft)]
    print("iteration", it, "max |regret| oracle", np.abs(orc.regret).max())
    print(" trunk regret err", np.abs(reg[:4] - orc.regret[:4]).max(), "max", np.abs(orc.regret[:4]).max(),
          "avg err", np.abs(avg[:4] - orc.avg[:4]).max())
    for i in range(1, st["n_local"]):
        par = st["parent"][i]
        n0 = st["node_base"][i] + st["node_k"][i]
        rows = np.array([ft.slot[n0 + j * st["node_m"][i]] for j in range(spec.boards.shape[0])])
        a, b = reg[rows], orc.regret[rows]
        x, y = avg[rows], orc.avg[rows]
        d = np.abs(a - b)
        w = np.unravel_index(d.argmax(), d.shape)
        dv = np.abs(x - y)
        wv = np.unravel_index(dv.argmax(), dv.shape)
        print(" child %2d (parent %2d seat %d): regret err %.3e (max %.3e) worst at board %d hand %d got %.6e ref %.6e | avg err %.3e n>1e-3: %d worst got %.4f ref %.4f regrets there got %s ref %s"
              % (i, par, st["kind"][par], d.max(), np.abs(b).max(), w[0], w[1], a[w], b[w], dv.max(), int((dv > 1e-3).sum()), x[wv], y[wv],
                 reg[rows[wv[0]] - (i - st["first_child"][par]): rows[wv[0]] - (i - st["first_child"][par]) + st["n_children"][par], wv[1]],
                 orc.regret[rows[wv[0]] - (i - st["first_child"][par]): rows[wv[0]] - (i - st["first_child"][par]) + st["n_children"][par], wv[1]]))

# ---- which seat-0 strategies differ on board 19 after the FIRST iteration?
spec = random_board_spec(48, 21)
orc = cfr2_c.Oracle2CSolver(ft, oracle_tree(ft).board_ranks, "CFRPlus", n_threads=8, lean=True)
s = BoardCFRSolver(g, args, spec)
# seat 0 half-iteration only
s._update_begin(0)
s._update_end(0)
import ctypes as C
t = C.byref(orc.t)
orc.L.orc2_values(t, orc.strat.ctypes.data, 1, 0)
orc.L.orc2_regret_update(t, 0, orc.algo, 0)
reg, avg = [x.cpu().numpy()[:, :ft.R].astype(np.float64) for x in s.natural_tables(ft)]
for b in (18, 19, 20):
    for d in (1, 2, 10):
        fc, A = st["first_child"][d], st["n_children"][d]
        n0 = st["node_base"][fc] + st["node_k"][fc]
        r0 = ft.slot[n0 + b * st["node_m"][fc]]
        a, o = reg[r0:r0 + A], orc.regret[r0:r0 + A]
        live = np.abs(o).sum(axis=0) + np.abs(a).sum(axis=0) > 0
        sa = np.where(a.sum(0) > 0, a / np.where(a.sum(0) > 0, a.sum(0), 1), 1.0 / A)
        so = np.where(o.sum(0) > 0, o / np.where(o.sum(0) > 0, o.sum(0), 1), 1.0 / A)
        dif = np.abs(sa - so).max(axis=0)
        bad = np.nonzero(dif > 1e-3)[0]
        print("board", b, "node", d, "regret abs err %.2e" % np.abs(a - o).max(), "hands with different strategy:", len(bad),
              "all-zero rows engine", int((a.sum(0) == 0).sum()), "oracle", int((o.sum(0) == 0).sum()))
        for h in bad[:4]:
            print("    hand", h, "engine", a[:, h], "oracle", o[:, h])
