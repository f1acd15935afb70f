"""torchrun helper: board-sharded Flop5Holdem CFR+ on WORLD_SIZE GPUs vs the same game on one GPU (rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pokerrl_b200.distributed import ShardedCFRSolver  # noqa: E402
from pokerrl_b200.game import games  # noqa: E402
from pokerrl_b200.game.holdem_boards import BoardSpec  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = games.Flop5Holdem
args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000], bet_sizes_list_as_frac_of_pot=[1.0])
deck = [0, 1, 2, 3, 4, 5, 6, 7, 48, 49, 50, 51]  # 792 boards -> 57 isomorphism classes
spec = BoardSpec.full_game(g.RULES, isomorphic=True, deck_subset=deck)
root_actions = None
algo = "CFRPlus"
if len(sys.argv) > 1 and sys.argv[1] == "hulh":  # multi-street Limit Hold'em sub-game, Linear CFR (BASELINE configs[3])
    from pokerrl_b200.game.holdem_boards import MultiStreetBoards
    g = games.LimitHoldem
    args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[48, 48], bet_sizes_list_as_frac_of_pot=[1.0])
    spec = MultiStreetBoards.subgame(g.RULES, (0, 5, 10), 2, 1, cards_per_layer=[[20, 21, 22, 23, 24], [30, 31, 32]])
    root_actions, algo = [1, 1], "LinearCFR"
s = ShardedCFRSolver(g, args, spec, algo, device="cuda:%d" % local, rank=rank, world=world, root_actions=root_actions)
trace = []
for _ in range(5):
    s.iteration(1)
    trace += [s.exploitability_current(), s.exploitability_average()]
single = None
if rank == 0:
    one = ShardedCFRSolver(g, args, spec, algo, device="cuda:0", rank=0, world=1, root_actions=root_actions)
    single = []
    for _ in range(5):
        one.iteration(1)
        single += [one.exploitability_current(), one.exploitability_average()]
gathered = [None] * world
dist.all_gather_object(gathered, trace)
if rank == 0:
    assert all(t == gathered[0] for t in gathered), "ranks disagree"
    print(json.dumps({"sharded": trace, "single": single, "world": world}))
dist.destroy_process_group()
