"""debug: one board, evaluation sweep, showdown-only / fold-only decomposition against the golden rows"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from gen_golden_twocard_common import make_reach  # noqa: E402
from pokerrl_b200 import _native as nat  # noqa: E402
from pokerrl_b200.board_engine import BoardCFRSolver  # noqa: E402
from pokerrl_b200.game import games  # noqa: E402
from pokerrl_b200.game.holdem_boards import BoardSpec  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "twocard_rows.npz"))
g = games.Flop5Holdem
args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000], bet_sizes_list_as_frac_of_pot=[1.0])
boards = GOLD["boards"][:4]
n = len(boards)
s = BoardCFRSolver(g, args, BoardSpec(boards, np.ones(n), np.ones(n), None, "dbg"))
hc = np.asarray(g.RULES.get_lut_holder().LUT_IDX_2_HOLE_CARDS).astype(np.int64)
reach = make_reach(int(GOLD["seed"]), GOLD["boards"], hc)
st = s.st
row = torch.zeros(s.ld, dtype=torch.float32, device=s.device)
pots = list(st["pot"])
for mode in ("both", "sd", "fold"):
    for i in range(st["n_local"]):
        k = st["kind"][i]
        s.g.pot[i] = pots[i] if (mode == "both" or (mode == "sd" and k == 4) or (mode == "fold" and k == 3)) else 0.0
    for p in (0, 1):
        csd = cf = 0.0
        for t in range(st["n_local"]):
            if st["kind"][t] < 3:
                continue
            w, i = 1.0, t
            while st["parent"][i] >= 0:
                w /= st["n_children"][st["parent"][i]]
                i = st["parent"][i]
            w *= s.g.pot[t] / 2
            if st["kind"][t] == 4:
                csd += w
            else:
                cf += -w if st["acted_last"][t] == p else w
        for b in range(2):
            row[:1326] = torch.from_numpy(reach[b]).to(s.device)
            s.t_mult.zero_()
            s.t_mult[b] = 1.0
            nat.call("prl_board_sweep", C.byref(s.g), p, 1, 0, 0, C.c_void_p(row.data_ptr()), 0, 0,
                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
            got = s.w_total[0].cpu().numpy().astype(np.float64) / 2.0 ** s.g.frac_bits
            ref = csd * GOLD["showdown"][b] + cf * GOLD["fold"][b]
            live = np.nonzero(ref != 0)[0]
            err = np.abs(got - ref).max() / np.abs(ref).max()
            ratio = got[live] / ref[live]
            print(mode, "p", p, "b", b, "err %.3e" % err, "ratio med %.4f min %.4f max %.4f" % (np.median(ratio), ratio.min(), ratio.max()),
                  "got[:4]", got[live[:4]], "ref[:4]", ref[live[:4]], "nonzero got", int((got != 0).sum()), "csd", csd, "cf", cf)
