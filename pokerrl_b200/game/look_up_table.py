"""Card / hand index look-up tables (host, built once; integer, bit-exact indexing contract).

Same tables, shapes, dtypes and accessor names as the reference's `LutHolderLeduc` / `LutHolderHoldem`
(`PokerRL/game/_/look_up_table.py:191-322`), whose Hold'em tables come from the binary-only `lib_luts.so`
(`cpp_wrappers/CppLUT.py:14-94`):
  card 1D <-> 2D     c = rank * N_SUITS + suit                     (test_look_up_table.py:110-113)
  range_idx          lexicographic index of (c1 < c2)              (test_look_up_table.py:136-143)
  LUT_HOLE_CARDS_2_IDX[c1, c2] = -2 where c1 >= c2                 (test_look_up_table.py:34-42)
Everything is computed with vectorised numpy instead of the reference's Python loops (~4 s for Hold'em there).
"""
from math import comb

import numpy as np

from pokerrl_b200.game.Poker import Poker

_ND = Poker.CARD_NOT_DEALT_TOKEN_1D


class LutHolder:
    def __init__(self, rules):
        self.rules = rules
        n_r, n_s, n_c, n_h = rules.N_RANKS, rules.N_SUITS, rules.N_CARDS_IN_DECK, rules.N_HOLE_CARDS
        cards = np.arange(n_c)
        self.LUT_1DCARD_2_2DCARD = np.stack([cards // n_s, cards % n_s], axis=1).astype(np.int8)
        self.LUT_2DCARD_2_1DCARD = cards.reshape(n_r, n_s).astype(np.int8)
        if n_h == 1:
            self.LUT_IDX_2_HOLE_CARDS = cards.reshape(-1, 1)
            self.LUT_HOLE_CARDS_2_IDX = cards.reshape(-1, 1)
            self.LUT_CARD_IN_WHAT_RANGE_IDXS = np.arange(rules.RANGE_SIZE).reshape(-1, 1)
        elif n_h == 2:
            c1, c2 = np.triu_indices(n_c, k=1)  # row-major => lexicographic (c1 < c2)
            self.LUT_IDX_2_HOLE_CARDS = np.stack([c1, c2], axis=1).astype(np.int8)
            h2i = np.full((n_c, n_c), -2, np.int16)
            h2i[c1, c2] = np.arange(c1.size, dtype=np.int16)
            self.LUT_HOLE_CARDS_2_IDX = h2i
            in_hand = (c1[None, :] == cards[:, None]) | (c2[None, :] == cards[:, None])  # [card, range_idx]
            self.LUT_CARD_IN_WHAT_RANGE_IDXS = np.nonzero(in_hand)[1].reshape(n_c, n_c - 1).astype(np.int32)
        else:
            raise NotImplementedError("games with > 2 hole cards")
        # private observation: per hole card one-hot rank (+ one-hot suit if suits matter)
        D = n_s + n_r
        obs = np.zeros((rules.RANGE_SIZE, D * n_h), np.float32)
        hc = self.LUT_IDX_2_HOLE_CARDS
        rows = np.arange(rules.RANGE_SIZE)
        for k in range(n_h):
            obs[rows, D * k + hc[:, k] // n_s] = 1
            if rules.SUITS_MATTER:
                obs[rows, D * k + n_r + hc[:, k] % n_s] = 1
        self.LUT_RANGE_IDX_TO_PRIVATE_OBS = obs
        rounds = rules.ALL_ROUNDS_LIST
        self.DICT_LUT_CARDS_DEALT_IN_TRANSITION_TO = {r: rules.n_cards_dealt_in_transition_to(r)
                                                      for r in (Poker.PREFLOP, Poker.FLOP, Poker.TURN, Poker.RIVER)}
        self.DICT_LUT_N_CARDS_OUT = {r: rules.n_cards_out_at(r)
                                     for r in (Poker.PREFLOP, Poker.FLOP, Poker.TURN, Poker.RIVER)}
        self.DICT_LUT_N_BOARDS = {r: comb(n_c, self.DICT_LUT_CARDS_DEALT_IN_TRANSITION_TO[r]) for r in rounds}
        self.DICT_LUT_N_BOARD_BRANCHES = {Poker.PREFLOP: 0}
        for r in rounds:
            if r != Poker.PREFLOP:
                free = n_c - self.DICT_LUT_N_CARDS_OUT[rules.ROUND_BEFORE[r]] - n_h
                self.DICT_LUT_N_BOARD_BRANCHES[r] = comb(free, self.DICT_LUT_CARDS_DEALT_IN_TRANSITION_TO[r])

    # ---- accessors (look_up_table.py:223-322)
    def get_1d_card(self, card_2d):
        if card_2d[0] == _ND:
            return _ND
        return self.LUT_2DCARD_2_1DCARD[card_2d[0], card_2d[1]]

    def get_1d_cards(self, cards_2d):
        cards_2d = np.asarray(cards_2d)
        if cards_2d.ndim == 0 or cards_2d.shape[0] == 0:
            return np.array([], dtype=np.int8)
        aa = np.where(cards_2d == _ND, 0, cards_2d)
        return np.where(cards_2d[:, 0] == _ND, _ND, self.LUT_2DCARD_2_1DCARD[aa[:, 0], aa[:, 1]])

    def get_2d_cards(self, cards_1d):
        cards_1d = np.asarray(cards_1d)
        if cards_1d.ndim == 0 or cards_1d.shape[0] == 0:
            return np.array([], dtype=np.int8)
        aa = np.where(cards_1d == _ND, 0, cards_1d)
        out = np.copy(self.LUT_1DCARD_2_2DCARD[aa]).reshape(-1, 2)
        out[cards_1d == _ND] = _ND
        return out

    def get_range_idx_from_hole_cards(self, hole_cards_2d):
        c = self.get_1d_cards(np.asarray(hole_cards_2d))
        if self.rules.N_HOLE_CARDS == 1:
            return self.LUT_HOLE_CARDS_2_IDX[c[0], 0]
        return self.LUT_HOLE_CARDS_2_IDX[min(c[0], c[1]), max(c[0], c[1])]

    def get_2d_hole_cards_from_range_idx(self, range_idx):
        return np.array([self.LUT_1DCARD_2_2DCARD[c] for c in self.LUT_IDX_2_HOLE_CARDS[range_idx]], dtype=np.int8)

    def get_1d_hole_cards_from_range_idx(self, range_idx):
        return np.copy(self.LUT_IDX_2_HOLE_CARDS[range_idx])


# names used by the reference
LutHolderLeduc = LutHolderHoldem = LutHolder
