"""The host tree compiler must reproduce the reference's PublicTree node for node
(PublicTree.py:111-293; fixtures: oracle/gen_golden_cfr.py:gen_tree)."""
import numpy as np
import pytest

from common import GAMES, golden, make_flat_tree


@pytest.mark.parametrize("name", list(GAMES))
def test_structure_matches_reference(name):
    ft = make_flat_tree(name)
    ref = golden("tree_%s.npz" % name)
    assert ft.n_nodes == len(ref["parent"]) == int(ref["n_nodes_reported"]) + 1
    assert ft.n_nonterm == int(ref["n_nonterm_reported"]) + 1  # the reference does not count the root
    perm = ft.dfs_permutation()
    assert sorted(ft.dfs.tolist()) == list(range(ft.n_nodes))
    for key, mine in (("kind", ft.kind), ("main_pot", ft.pot), ("n_children", ft.n_children),
                      ("round", ft.round), ("acted_last", ft.acted_last), ("action", ft.action),
                      ("stack", ft.stack), ("bet", ft.bet), ("board", ft.node_board_cards())):
        assert np.array_equal(mine[perm], ref[key]), key
    par = ft.parent[perm]
    assert np.array_equal(np.where(par >= 0, ft.dfs[np.maximum(par, 0)], -1), ref["parent"])


@pytest.mark.parametrize("name", list(GAMES))
def test_layout_invariants(name):
    ft = make_flat_tree(name)
    N = ft.n_nodes
    # depth-sorted, children contiguous and one level below the parent
    lvl = np.searchsorted(ft.level_start, np.arange(N), side="right") - 1
    nz = ft.parent >= 0
    assert np.all(lvl[nz] == lvl[ft.parent[nz]] + 1)
    has = ft.first_child >= 0
    idx = np.nonzero(has)[0]
    for n in idx[:: max(1, len(idx) // 500)]:
        ch = np.arange(ft.first_child[n], ft.first_child[n] + ft.n_children[n])
        assert np.all(ft.parent[ch] == n)
    # slots: one per child of a decision node, contiguous per decision node
    dec = (ft.kind <= 1) & has
    assert ft.n_slots == int(ft.n_children[dec].sum())
    assert np.all(ft.first_slot[dec] == ft.slot[ft.first_child[dec]])
    assert np.all(ft.slot[ft.first_child[dec] + ft.n_children[dec] - 1] == ft.first_slot[dec] + ft.n_children[dec] - 1)
