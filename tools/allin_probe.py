"""Timing probe of the all-in equity path on one GPU: python tools/allin_probe.py [n_cols]
Full-game equity matrix build (134 459 isomorphism classes) and the tensor-core product, CUDA events."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pokerrl_b200.allin import AllinEquity  # noqa: E402
from pokerrl_b200.game import games  # noqa: E402
from pokerrl_b200.game.holdem_boards import BoardSpec  # noqa: E402

n_cols = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rules = games.Flop5Holdem.RULES
spec = BoardSpec.full_game(rules)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eq = AllinEquity(rules, spec)
e1.record()
torch.cuda.synchronize()
build_ms = e0.elapsed_time(e1)
x = torch.rand(n_cols, 1328, device="cuda") ** 3
y = eq.values(x)
torch.cuda.synchronize()
ts = []
for _ in range(50):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eq.values(x, out=y)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
flops = 2.0 * 1326 * 1326 * n_cols  # useful fp32-equivalent flops; the tensor cores execute 6 bf16 plane products of padded tiles
issued = 6 * 2.0 * 1408 * 1344 * 16
print(json.dumps({"boards": int(spec.boards.shape[0]), "equity_build_ms": build_ms, "n_cols": n_cols, "values_ms_median": ms,
                  "values_ms_min": float(min(ts)), "operand_bytes": int(eq.tiles.numel()),
                  "useful_gflops_per_s": flops / ms / 1e6, "issued_bf16_tflops_per_s": issued / ms / 1e9,
                  "note": "two launches per call (allin_gemm_kernel 11 x 21 CTAs + allin_finish_kernel); latency-bound by design"}))
