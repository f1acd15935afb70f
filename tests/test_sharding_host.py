"""Host-side logic of the multi-GPU path (no GPU): board shards partition the board set, weights are preserved, and a
world_size-2 gloo all-reduce over the chance-sum layout reproduces the unsharded sum."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pokerrl_b200.game.games import FlopHoldemRules, HoldemRules
from pokerrl_b200.game.holdem_boards import BoardSpec, MultiStreetBoards


def test_board_shards_partition_the_spec():
    from pokerrl_b200.distributed import shard_board_spec
    deck = [0, 1, 2, 3, 4, 5, 6, 7, 48, 49, 50, 51]
    spec = BoardSpec.full_game(FlopHoldemRules, isomorphic=True, deck_subset=deck)
    for world in (2, 3, 8):
        shards = [shard_board_spec(spec, r, world) for r in range(world)]
        allb = np.concatenate([s.boards for s in shards])
        assert sorted(map(tuple, allb.tolist())) == sorted(map(tuple, spec.boards.tolist()))
        assert np.isclose(sum(s.board_mult.sum() for s in shards), spec.board_mult.sum())
        assert max(s.boards.shape[0] for s in shards) - min(s.boards.shape[0] for s in shards) <= 1


def test_multi_street_shards_follow_their_parents():
    from pokerrl_b200.distributed import shard_board_spec
    spec = MultiStreetBoards.subgame(HoldemRules, (0, 5, 10), 2, 1, cards_per_layer=[list(range(20, 27)), [30, 31, 32]])
    shards = [shard_board_spec(spec, r, 2) for r in range(2)]
    turns = np.concatenate([s.boards[1] for s in shards])
    assert sorted(map(tuple, turns.tolist())) == sorted(map(tuple, spec.boards[1].tolist()))
    for s in shards:
        assert s.boards[2].shape[0] == 3 * s.boards[1].shape[0]
        # every river board extends its (re-indexed) parent turn board
        assert np.array_equal(s.boards[2][:, :4], s.boards[1][s.parents[2]])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ld, n_chance, n_boundary, chunks = 8, 3, 2, 2
    rng = np.random.default_rng(rank)
    ws = torch.from_numpy(rng.random(4 * n_chance * (chunks + 1) * ld).astype(np.float32))
    local = ws.clone()
    w_off = 4 * n_chance * chunks * ld
    for arr in (0, 2):  # ev of both seats
        o = w_off + arr * n_chance * ld
        dist.all_reduce(ws[o:o + n_boundary * ld], op=dist.ReduceOp.SUM)
    torch.save((local, ws), os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_gloo_allreduce_of_boundary_chance_sums(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, 29541, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    ld, n_chance, n_boundary, chunks = 8, 3, 2, 2
    w_off = 4 * n_chance * chunks * ld
    for r in range(world):
        local, reduced = res[r]
        expect = local.clone()
        for arr in (0, 2):
            o = w_off + arr * n_chance * ld
            expect[o:o + n_boundary * ld] = sum(res[q][0][o:o + n_boundary * ld] for q in range(world))
        assert torch.equal(reduced, expect)  # only the boundary slices changed, and they hold the sum over ranks
