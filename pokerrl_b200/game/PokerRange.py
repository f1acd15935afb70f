"""Distribution over a player's private hands, tracked along a hand (`PokerRL/game/PokerRange.py:9-160`): the LBR evaluator
multiplies it by the agent's action probabilities after every agent action, removes the hands the board blocks when a
street is dealt and renormalises (a range that sums to 0 falls back to uniform, PokerRange.py:42-47).  float32 like the
reference; the blocker masks are one vectorised table instead of per-card slices."""
import numpy as np

from pokerrl_b200.game.Poker import Poker


class PokerRange:
    def __init__(self, env_bldr):
        rules = env_bldr.rules
        assert rules.N_HOLE_CARDS <= 2
        self._env_bldr, self._R = env_bldr, rules.RANGE_SIZE
        lut = env_bldr.lut_holder
        hc = np.asarray(lut.LUT_IDX_2_HOLE_CARDS).reshape(self._R, rules.N_HOLE_CARDS).astype(np.int64)
        self._holds = np.zeros((rules.N_CARDS_IN_DECK, self._R), bool)  # [card, hand]: the hand contains the card
        for k in range(rules.N_HOLE_CARDS):
            self._holds[hc[:, k], np.arange(self._R)] = True
        self._range = None
        self.reset()

    range = property(lambda self: self._range)

    def get_range(self):
        return np.copy(self._range)

    def get_card_probs(self):
        return (self._holds.astype(np.float32) @ self._range).astype(np.float32)

    def reset(self):
        self._range = np.full(self._R, 1.0 / self._R, dtype=np.float32)

    def normalize(self):
        total = np.sum(self._range, axis=-1)
        if total == 0:
            self.reset()
        else:
            self._range = self._range / total

    def mul_and_norm(self, mul_vector):
        self._range *= mul_vector
        self.normalize()

    def update_after_action(self, action, all_a_probs_for_all_hands):
        self._range *= all_a_probs_for_all_hands[:, action]
        self.normalize()

    def set_cards_to_zero_prob(self, cards_2d):
        cards = np.asarray(self._env_bldr.lut_holder.get_1d_cards(cards_2d=np.asarray(cards_2d))).reshape(-1)
        cards = cards[cards != Poker.CARD_NOT_DEALT_TOKEN_1D]
        if cards.size:
            self._range[self._holds[cards].any(axis=0)] = 0
        self.normalize()

    def update_after_new_round(self, new_round, board_now_2d):
        lut, rules = self._env_bldr.lut_holder, self._env_bldr.rules
        n_now, n_before = lut.DICT_LUT_N_CARDS_OUT[new_round], lut.DICT_LUT_N_CARDS_OUT[rules.ROUND_BEFORE[new_round]]
        self.set_cards_to_zero_prob(cards_2d=np.asarray(board_now_2d)[n_before:n_now].reshape(-1, 2))

    def state_dict(self):
        return {"range": np.copy(self._range)}

    def load_state_dict(self, state):
        self._range = np.copy(state["range"])
