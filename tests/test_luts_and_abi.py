"""CPU-side tests: LUT tables against the reference's, the hand-eval oracle against the golden ranks, and that the
C-ABI library loads and exports every symbol include/pokerrl_b200.h declares (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import golden
from pokerrl_b200.game import games

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,game", [("holdem", games.DiscretizedNLHoldem), ("leduc", games.DiscretizedNLLeduc)])
def test_luts_match_reference(name, game):
    g = golden("luts.npz")
    L = game.get_lut_holder()
    for k in ("LUT_1DCARD_2_2DCARD", "LUT_2DCARD_2_1DCARD", "LUT_IDX_2_HOLE_CARDS", "LUT_HOLE_CARDS_2_IDX",
              "LUT_CARD_IN_WHAT_RANGE_IDXS", "LUT_RANGE_IDX_TO_PRIVATE_OBS"):
        ref = g["%s_%s" % (name, k)]
        mine = getattr(L, k)
        assert np.array_equal(mine, ref), k
        if name == "holdem":
            assert mine.dtype == ref.dtype, (k, mine.dtype, ref.dtype)  # test_look_up_table.py:16-56 pins dtypes
    for k in ("DICT_LUT_N_BOARDS", "DICT_LUT_N_CARDS_OUT", "DICT_LUT_CARDS_DEALT_IN_TRANSITION_TO",
              "DICT_LUT_N_BOARD_BRANCHES"):
        ref = {int(a): int(b) for a, b in g["%s_%s" % (name, k)]}
        mine = getattr(L, k)
        for kk, vv in ref.items():
            if kk in game.RULES.ALL_ROUNDS_LIST or k in ("DICT_LUT_N_CARDS_OUT", "DICT_LUT_CARDS_DEALT_IN_TRANSITION_TO"):
                assert mine[kk] == vv, (k, kk)


def test_lut_accessors_reference_tests():
    """value checks of the reference's own test_look_up_table.py:110-167"""
    L = games.DiscretizedNLHoldem.get_lut_holder()
    assert L.get_1d_card(np.array([0, 3], np.int8)) == 3 and L.get_1d_card(np.array([12, 3], np.int8)) == 51
    n = 0
    for c1 in range(52):
        for c2 in range(c1 + 1, 52):
            assert L.LUT_HOLE_CARDS_2_IDX[c1, c2] == n
            n += 1
    assert np.array_equal(L.get_2d_cards(np.array([5, -127], np.int8)), np.array([[1, 1], [-127, -127]], np.int8))
    assert L.get_range_idx_from_hole_cards(np.array([[12, 3], [0, 0]], np.int8)) == L.LUT_HOLE_CARDS_2_IDX[0, 51]


def test_hand_eval_oracle_matches_reference_binary_fixture():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    orc = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libhand_eval_oracle.so"))
    orc.orc_rank_boards.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    g = golden("hand_ranks.npz")
    boards = np.ascontiguousarray(g["boards"])
    out = np.zeros((len(boards), 1326), np.int32)
    orc.orc_rank_boards(out.ctypes.data, boards.ctypes.data, len(boards))
    assert np.array_equal(out, g["ranks"])


def test_c_abi_library_exports_every_declared_symbol():
    from pokerrl_b200 import _native
    L = _native.lib()
    hdr = open(os.path.join(ROOT, "include", "pokerrl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b((?:prl_|get_)[a-z0-9_]+)\s*\(", hdr)
    assert len(names) >= 15
    for n in sorted(set(names)):
        assert hasattr(L, n), n
    assert L.prl_abi_version() == _native.ABI_VERSION


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct in include/pokerrl_b200.h, as gcc lays them out, against the ctypes mirrors
    in pokerrl_b200/_native.py (a silent mismatch would shift every pointer the kernels read)"""
    import ctypes as C
    import subprocess
    from pokerrl_b200 import _native
    pairs = [("prl_tree_t", _native.PrlTree), ("prl_buffers_t", _native.PrlBuffers), ("prl_env_cfg_t", _native.PrlEnvCfg), ("prl_board_game_t", _native.PrlBoardGame), ("prl_trunk_t", _native.PrlTrunk)]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "pokerrl_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src.append('printf("abi %d\\n", PRL_ABI_VERSION); return 0; }')
    c_file, exe = tmp_path / "layout.c", tmp_path / "layout"
    c_file.write_text("\n".join(src))
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(c_file), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(got["abi"]) == _native.ABI_VERSION
    for cname, cls in pairs:
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_reference_binding_sequence_against_this_library(golden_dir):
    """Replays what the reference does with its native libraries - CppWrapper (PokerRL/_/CppWrapper.py:10-27: LoadLibrary, 2-D
    arrays as arrays of row pointers), CppLibHoldemLuts.__init__ + getters (CppLUT.py:16-94) and CppHandeval.__init__
    (CppHandeval.py:19-33) - against libpokerrl_b200.so instead of lib_luts.so / lib_hand_eval.so: every symbol the
    reference sets argtypes on exists, and the LUT natives (host code, no GPU needed) fill the reference's buffer shapes with
    the reference's tables (tests/golden/luts.npz)."""
    import ctypes
    from pokerrl_b200 import _native
    lib = ctypes.cdll.LoadLibrary(_native.LIB_PATH)
    ARR_2D = np.ctypeslib.ndpointer(dtype=np.intp, ndim=1, flags="C")

    def rows(arr):  # CppWrapper.np_2d_arr_to_c
        return (arr.__array_interface__["data"][0] + np.arange(arr.shape[0]) * arr.strides[0]).astype(np.intp)

    for f in ("get_hole_card_2_idx_lut", "get_idx_2_hole_card_lut", "get_idx_2_flop_lut", "get_idx_2_turn_lut",
              "get_idx_2_river_lut"):  # CppLUT.py:22-35
        getattr(lib, f).argtypes = [ARR_2D]
        getattr(lib, f).restype = None
    lib.get_hand_rank_52_holdem.argtypes = [ARR_2D, ARR_2D]  # CppHandeval.py:22-33
    lib.get_hand_rank_52_holdem.restype = ctypes.c_int32
    lib.get_hand_rank_all_hands_on_given_boards_52_holdem.argtypes = [ARR_2D, ARR_2D, ctypes.c_int32, ARR_2D, ARR_2D]
    lib.get_hand_rank_all_hands_on_given_boards_52_holdem.restype = None
    gold = np.load(os.path.join(golden_dir, "luts.npz"))
    n_boards = dict(gold["holdem_DICT_LUT_N_BOARDS"].tolist())
    n_out = dict(gold["holdem_DICT_LUT_N_CARDS_OUT"].tolist())
    a = np.full((1326, 2), -2, np.int8)  # CppLUT.py:38-41
    lib.get_idx_2_hole_card_lut(rows(a))
    assert np.array_equal(a, gold["holdem_LUT_IDX_2_HOLE_CARDS"])
    b = np.full((52, 52), -2, np.int16)  # CppLUT.py:43-46
    lib.get_hole_card_2_idx_lut(rows(b))
    assert np.array_equal(b, gold["holdem_LUT_HOLE_CARDS_2_IDX"])
    from itertools import combinations
    flop = np.full((n_boards[1], n_out[1]), -2, np.int8)  # CppLUT.py:47-54: [22100][3]
    lib.get_idx_2_flop_lut(rows(flop))
    assert np.array_equal(flop, np.array(list(combinations(range(52), 3)), np.int8))
    for f, rnd in (("get_idx_2_turn_lut", 2), ("get_idx_2_river_lut", 3)):  # [52][4], [52][5]: in-bounds, card per row
        t = np.full((n_boards[rnd], n_out[rnd]), -2, np.int8)
        getattr(lib, f)(rows(t))
        assert np.array_equal(t[:, 0], np.arange(52)) and np.all(t[:, 1:] == -2)
    lib.get_1d_card.argtypes, lib.get_1d_card.restype = [ctypes.c_void_p], ctypes.c_int8
    lib.get_2d_card.argtypes, lib.get_2d_card.restype = [ctypes.c_int8, ctypes.c_void_p], None
    for c in range(52):  # look_up_table.py:102-119
        out = np.empty(2, np.int8)
        lib.get_2d_card(c, out.ctypes.data)
        assert np.array_equal(out, gold["holdem_LUT_1DCARD_2_2DCARD"][c])
        assert lib.get_1d_card(out.ctypes.data) == gold["holdem_LUT_2DCARD_2_1DCARD"][out[0], out[1]] == c
