"""Host-side structure of the two-card trees (no GPU): Flop5Holdem betting structure of SURVEY.md §8 and the
suit-isomorphism board classes."""
import numpy as np

from pokerrl_b200.game import games, holdem_boards as hb
from pokerrl_b200.game.flat_tree import enumerate_betting_tree
from pokerrl_b200.game.hu_engine import HUBetting
from twocard_common import fhp_tree, random_board_spec


def test_flop5holdem_betting_structure():
    g = games.Flop5Holdem
    args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000], bet_sizes_list_as_frac_of_pot=[1.0])
    A = enumerate_betting_tree(HUBetting(g, args))
    post = [n for n in A if n.cdepth == 1]
    dec = [n for n in post if n.kind <= 1]
    assert len(dec) == 6 and sum(len(n.children) for n in dec) == 14          # SURVEY.md §8 header
    assert sum(n.kind == 4 for n in post) == 5 and sum(n.kind == 3 for n in post) == 4
    assert [len(n.children) for n in A if n.cdepth == 0 and n.kind <= 1] == [2, 2]  # SB {fold, raise}, BB {fold, call}


def test_limit_holdem_betting_structure():
    g = games.LimitHoldem
    args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[48, 48], bet_sizes_list_as_frac_of_pot=[1.0])
    A = enumerate_betting_tree(HUBetting(g, args))
    assert sum(n.kind == 3 for n in A) == 5103 and sum(n.kind == 4 for n in A) == 5103  # 10 206 terminal sequences
    assert sum(n.kind <= 1 and n.children != [] for n in A) == 6378


def test_suit_isomorphism_classes_small_deck():
    deck = [0, 1, 2, 3, 4, 5, 6, 7, 48, 49, 50, 51]  # ranks 2, 3, A in all four suits: closed under suit permutations
    boards = hb.all_boards(deck, 5)
    reps, orbit = hb.canonical_boards(boards)
    assert orbit.sum() == len(boards) == 792
    perms = hb.suit_permutation_hand_tables()
    assert perms.shape == (24, 1326) and np.array_equal(np.sort(perms, axis=1), np.tile(np.arange(1326), (24, 1)))
    # every board is a suit permutation of exactly one representative
    keys = {tuple(r) for r in reps.tolist()}
    assert len(keys) == len(reps)


def test_flat_tree_layout_two_card():
    ft = fhp_tree(random_board_spec(20, 0))
    nb = ft.board_spec.boards.shape[0]
    assert ft.n_nodes == 5 + 15 * nb and ft.n_slots == 4 + 14 * nb
    ch = np.nonzero(ft.kind == 2)[0]
    assert len(ch) == 1 and ft.n_children[ch[0]] == nb
    kids = np.arange(ft.first_child[ch[0]], ft.first_child[ch[0]] + nb)
    assert np.array_equal(ft.board[kids], 1 + np.arange(nb))  # global board ids, 0 = the empty board


def test_structure_records_restate_the_pointer_chains():
    """prl_tree_t.node_rec2 / work_rec2 (pokerrl_b200/solver.py:structure_records) hold exactly what the v1 kernels read
    through parent -> first_child -> slot chains"""
    from pokerrl_b200.solver import structure_records
    from twocard_common import fhp_tree, random_board_spec
    ft = fhp_tree(random_board_spec(5, 3))
    order, _ = ft.work_order()
    nrec, wrec = structure_records(ft, order)
    for n in range(ft.n_nodes):
        p = int(ft.parent[n])
        assert nrec[n, 0] == p and nrec[n, 1] == ft.slot[n]
        if p >= 0:
            assert nrec[n, 2] == ft.slot[ft.first_child[p]]
            assert nrec[n, 3] & 0xff == ft.kind[p] and nrec[n, 3] >> 8 == ft.n_children[p]
    for t in range(ft.n_nodes):
        n = int(order[t])
        assert wrec[t, 0] == n and wrec[t, 3] & 0xff == ft.kind[n]
        if ft.n_children[n] > 0:
            assert wrec[t, 3] >> 8 == ft.n_children[n]
            assert wrec[t, 1] == ft.first_child[n] and wrec[t, 2] == ft.slot[ft.first_child[n]]
        else:  # terminal entries carry what terminal2_kernel_v3 needs
            assert wrec[t, 1] == ft.board[n] and wrec[t, 3] >> 8 == (int(ft.acted_last[n]) & 0xff)
            assert wrec[t, 2:3].view(np.float32)[0] == np.float32(ft.pot[n])
