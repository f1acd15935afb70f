"""Helpers for the two-card (Hold'em family) parity tests."""
import ctypes as C
import os

import numpy as np

import cfr2_numpy as o2
from pokerrl_b200.game import games
from pokerrl_b200.game.flat_tree import FlatTree
from pokerrl_b200.game.holdem_boards import BoardSpec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fhp_tree(board_spec, stack=20000):
    g = games.Flop5Holdem
    args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[stack, stack], bet_sizes_list_as_frac_of_pot=[1.0])
    return FlatTree(g, args, board_spec=board_spec)


def oracle_ranks(boards):
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    orc = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libhand_eval_oracle.so"))
    orc.orc_rank_boards.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    boards = np.ascontiguousarray(boards, np.int8)
    out = np.zeros((len(boards), 1326), np.int32)
    orc.orc_rank_boards(out.ctypes.data, boards.ctypes.data, len(boards))
    return out


def oracle_tree(ft):
    """float64 oracle over the same flat tree / board spec"""
    spec = ft.board_spec
    lut = ft.rules.get_lut_holder()
    bc = ft.board_cards()  # every board of every chance depth, global board id order
    ranks = np.full((bc.shape[0], ft.R), -1, np.int32)
    complete = np.nonzero((bc >= 0).sum(axis=1) == 5)[0]
    if complete.size:
        ranks[complete] = oracle_ranks(bc[complete])
    t = o2.Oracle2Tree(ft, lut.LUT_IDX_2_HOLE_CARDS, ranks, ft.board_prob, ft.board_mult, spec.sym_perm)
    if getattr(ft, "allin_spec", None) is not None:
        hc = np.asarray(lut.LUT_IDX_2_HOLE_CARDS).astype(np.int64)
        t.allin_equity = {key: o2.allin_equity_matrix(oracle_ranks(boards), w, hc, ft.rules.N_CARDS_IN_DECK, sym)
                          for key, (boards, w, sym) in ft.allin_completions().items()}
    return t


def oracle_allin_equity(rules, spec, ranks=None):
    """float64 equity matrix of the boards of `spec` (hand strengths from the C evaluator pinned to lib_hand_eval.so)"""
    rk = oracle_ranks(spec.boards) if ranks is None else ranks
    w = np.asarray(spec.board_prob, np.float64) * np.asarray(spec.board_mult, np.float64)
    hc = np.asarray(rules.get_lut_holder().LUT_IDX_2_HOLE_CARDS).astype(np.int64)
    return o2.allin_equity_matrix(rk, w, hc, rules.N_CARDS_IN_DECK, spec.sym_perm)


def hulh_flop_subgame(cards_per_layer, root_board=(0, 5, 10), stack=48):
    """LimitHoldem sub-game rooted at a flop after SB limps / BB checks; turn and river cards restricted for tests"""
    from pokerrl_b200.game.holdem_boards import MultiStreetBoards
    g = games.LimitHoldem
    args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[stack, stack], bet_sizes_list_as_frac_of_pot=[1.0])
    spec = MultiStreetBoards.subgame(g.RULES, root_board, 2, 1, cards_per_layer=cards_per_layer)
    return FlatTree(g, args, board_spec=spec, root_actions=[1, 1])


def random_board_spec(n, seed):
    rng = np.random.default_rng(seed)
    boards = np.unique(np.sort(np.stack([rng.choice(52, 5, replace=False) for _ in range(n)]), axis=1), axis=0)
    return BoardSpec(boards.astype(np.int8), np.full(len(boards), 1.0 / len(boards)), np.ones(len(boards)), None,
                     "%d random boards" % len(boards))


def nl_flop_subgame(stack=600, cards_per_layer=((20, 25, 31), (40, 44, 49, 51)), root_board=(0, 5, 10)):
    """DiscretizedNLHoldem (pot-size bets) rooted at a flop after SB calls / BB checks, short stacks: all-in showdowns
    happen on the flop (turn + river to come) and on the turn (river to come); turn / river cards restricted for tests"""
    from pokerrl_b200.game.holdem_boards import MultiStreetBoards
    g = games.DiscretizedNLHoldem
    args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[stack, stack], bet_sizes_list_as_frac_of_pot=[1.0])
    spec = MultiStreetBoards.subgame(g.RULES, root_board, 2, 1, cards_per_layer=[list(c) for c in cards_per_layer])
    return FlatTree(g, args, board_spec=spec, root_actions=[1, 1])
