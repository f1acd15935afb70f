"""PokerViz export: pokerrl_b200.game.tree_export against `PublicTree.get_tree_as_dict()` of the
reference (tests/golden/export_*.json.gz, written by oracle/gen_golden_export.py from the reference itself).  The node
vectors fed in are the reference's own (tests/golden/values_*.npz) - on the GPU they come from the device buffers, which
tests/test_gpu_cfr.py pins bit-for-bit to the same fixtures."""
import gzip
import json
import os

import numpy as np
import pytest

from common import GOLDEN, golden, make_flat_tree
from pokerrl_b200.game.tree_export import cards2str, export_tree_dict, write_tree_js


def _load(game, state):
    return json.loads(gzip.open(os.path.join(GOLDEN, "export_%s_%s.json.gz" % (game, state))).read())


def _first_diff(a, b, path="root"):
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys(), (path, list(a), list(b))
        for k in a:
            d = _first_diff(a[k], b[k], path + "." + str(k))
            if d:
                return d
        return None
    if isinstance(a, list):
        if len(a) != len(b):
            return (path, len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            d = _first_diff(x, y, "%s[%d]" % (path, i))
            if d:
                return d
        return None
    return None if a == b else (path, a, b)


def _chance_strategy(ft):
    bc = ft.node_board_cards()

    def strategy_of(n):
        fc, A = int(ft.first_child[n]), int(ft.n_children[n])
        if ft.kind[n] == 2:  # StrategyFiller.py:148-169
            s = np.zeros((ft.R, A), np.float32)
            for c in range(A):
                mask = np.ones(ft.R, bool)
                mask[bc[fc + c][bc[fc + c] >= 0]] = False
                s[mask, c] = 1.0 / (ft.rules.N_CARDS_IN_DECK - 2)
            return s
        return None
    return strategy_of


@pytest.mark.parametrize("game", ["StandardLeduc", "NLLeduc_POT"])
def test_export_of_a_freshly_built_tree(game):
    ft = make_flat_tree(game)
    want = _load(game, "built")
    assert _first_diff(export_tree_dict(ft), want) is None


@pytest.mark.parametrize("game", ["StandardLeduc", "NLLeduc_POT"])
def test_export_with_uniform_profile_values(game, tmp_path):
    ft = make_flat_tree(game)
    g = golden("values_%s.npz" % game)
    flat = {k: np.ascontiguousarray(g["uniform_" + k][ft.dfs].transpose(1, 0, 2)) for k in ("reach", "ev", "ev_br")}
    chance = _chance_strategy(ft)

    def strategy_of(n):
        if ft.kind[n] == 2:
            return chance(n)
        return np.full((ft.R, int(ft.n_children[n])), 1.0 / float(ft.n_children[n]))  # StrategyFiller.py:61-62
    got = export_tree_dict(ft, flat["reach"], flat["ev"], flat["ev_br"], strategy_of)
    assert _first_diff(got, _load(game, "uniform")) is None
    write_tree_js(str(tmp_path / "data.js"), got)
    txt = open(tmp_path / "data.js").read()
    assert txt.startswith("const data=") and json.loads(txt[len("const data="):]) == got


def test_card_strings():
    from pokerrl_b200.game import games
    assert cards2str(games.StandardLeduc.RULES, [0, -127]) == "2a, "
    assert cards2str(games.Flop5Holdem.RULES, [0, 5, 10, 51, 35]) == "2h, 3d, 4s, Ac, Tc, "
