"""torchrun helper: board engine on WORLD_SIZE GPUs (boards round-robin, NCCL all-reduce of the int64 chance sums) against
the same game on one GPU (rank 0): exploitability traces and trunk tables must be IDENTICAL (integer sums are exact)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pokerrl_b200.board_engine import BoardCFRSolver  # noqa: E402
from pokerrl_b200.game import games  # noqa: E402
from pokerrl_b200.game.holdem_boards import BoardSpec  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = games.Flop5Holdem
args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000], bet_sizes_list_as_frac_of_pot=[1.0])
full = BoardSpec.full_game(g.RULES)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 999
algo = sys.argv[2] if len(sys.argv) > 2 else "CFRPlus"
spec = BoardSpec(full.boards[:n], full.board_prob[:n], full.board_mult[:n], full.sym_perm, "first %d" % n)
s = BoardCFRSolver(g, args, spec, algo=algo, device="cuda:%d" % local, rank=rank, world=world)
trace = []
for _ in range(6):
    s.iteration(1)
    trace += [s.exploitability_current(), s.exploitability_average()]
chk = [float(s.bufs.regret.double().sum()), float(s.bufs.avg.double().sum())]
single = single_chk = None
if rank == 0:
    one = BoardCFRSolver(g, args, spec, algo=algo, device="cuda:0")
    single = []
    for _ in range(6):
        one.iteration(1)
        single += [one.exploitability_current(), one.exploitability_average()]
    single_chk = [float(one.bufs.regret.double().sum()), float(one.bufs.avg.double().sum())]
gathered = [None] * world
dist.all_gather_object(gathered, (trace, chk))
if rank == 0:
    assert all(t == gathered[0] for t in gathered), "ranks disagree"
    print(json.dumps({"sharded": trace, "single": single, "checksums": chk, "single_checksums": single_chk, "world": world}))
dist.destroy_process_group()
