"""GPU hand evaluator through the C ABI: bit-exact against the reference binary's golden ranks and the C oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from common import golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_boards_match_reference_binary():
    from pokerrl_b200.hand_eval import hand_rank_all_hands_on_given_boards
    g = golden("hand_ranks.npz")
    out = hand_rank_all_hands_on_given_boards(g["boards"]).cpu().numpy()
    assert np.array_equal(out, g["ranks"])


def test_random_boards_and_hands_match_oracle():
    from pokerrl_b200.hand_eval import hand_rank_7, hand_rank_all_hands_on_given_boards
    orc = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libhand_eval_oracle.so"))
    orc.orc_rank_boards.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    orc.orc_rank7_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    rng = np.random.default_rng(7)
    boards = np.stack([rng.choice(52, 5, replace=False) for _ in range(4000)]).astype(np.int8)
    ref = np.zeros((len(boards), 1326), np.int32)
    orc.orc_rank_boards(ref.ctypes.data, boards.ctypes.data, len(boards))
    assert np.array_equal(hand_rank_all_hands_on_given_boards(boards).cpu().numpy(), ref)
    cards = np.stack([rng.choice(52, 7, replace=False) for _ in range(200000)]).astype(np.int8)
    ref7 = np.zeros(len(cards), np.int32)
    orc.orc_rank7_batch(ref7.ctypes.data, cards.ctypes.data, len(cards))
    assert np.array_equal(hand_rank_7(cards).cpu().numpy(), ref7)


def test_legacy_native_signatures():
    """The reference's own ctypes wrappers' calling convention (CppWrapper.py:24-27) against this library."""
    from pokerrl_b200 import _native
    from pokerrl_b200.game.games import DiscretizedNLHoldem
    L = _native.lib()
    lut = DiscretizedNLHoldem.get_lut_holder()
    g = golden("hand_ranks.npz")

    def rows(a):
        return (a.__array_interface__['data'][0] + np.arange(a.shape[0]) * a.strides[0]).astype(np.intp)

    argt = np.ctypeslib.ndpointer(dtype=np.intp, ndim=1, flags='C')
    L.get_hand_rank_all_hands_on_given_boards_52_holdem.argtypes = [argt, argt, C.c_int32, argt, argt]
    L.get_hand_rank_all_hands_on_given_boards_52_holdem.restype = None
    boards = np.ascontiguousarray(g["boards"][:50])
    out = np.full((50, 1326), -1, np.int32)
    idx2hc = np.ascontiguousarray(lut.LUT_IDX_2_HOLE_CARDS)
    c2d = np.ascontiguousarray(lut.LUT_1DCARD_2_2DCARD)
    L.get_hand_rank_all_hands_on_given_boards_52_holdem(rows(out), rows(boards), 50, rows(idx2hc), rows(c2d))
    assert np.array_equal(out, g["ranks"][:50])
    L.get_hand_rank_52_holdem.argtypes = [argt, argt]
    L.get_hand_rank_52_holdem.restype = C.c_int32
    for b in range(5):
        for h in (0, 700, 1325):
            if g["ranks"][b, h] < 0:
                continue
            hand2d = np.ascontiguousarray(lut.get_2d_cards(lut.LUT_IDX_2_HOLE_CARDS[h]))
            board2d = np.ascontiguousarray(lut.get_2d_cards(boards[b]))
            assert L.get_hand_rank_52_holdem(rows(hand2d), rows(board2d)) == g["ranks"][b, h]
    h2i = np.full((52, 52), -2, np.int16)
    L.get_hole_card_2_idx_lut.argtypes = [argt]
    L.get_hole_card_2_idx_lut(rows(h2i))
    assert np.array_equal(h2i, lut.LUT_HOLE_CARDS_2_IDX)
