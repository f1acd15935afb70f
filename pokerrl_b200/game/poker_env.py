"""Single-table `PokerEnv` with the reference's call surface - `reset(deck_state_dict=None)`, `step(action)`,
`state_dict()`, `load_state_dict()`, `get_legal_actions()`, `cards_state_dict()` (PokerRL/game/_/rl_env/base/PokerEnv.py:
1075-1122, 1148-1159, 1161-1251, 1313-1330, 1361-1371) - on top of the batched device engine (`BatchedPokerEnv` with one
table; csrc/env_kernels.cu does all the game logic).  Heads-up limit / discretized games, evaluation mode (no stack
randomisation): the configuration the tabular CFR path and BASELINE.json configs[4] use.

`step` returns the reference's tuple `(obs float32[obs_size], rewards float[2], done bool, info)`; `info` is the
reference's `[False, None]` placeholder (the tree builder's pre-transition snapshots are produced by the flat-tree
compiler, not by stepping an env).  `state_dict()` carries the reference's public keys (PokerEnvStateDictEnums.py:9-36)
for inspection plus the raw device state, which is what `load_state_dict` restores exactly.
"""
import numpy as np
import torch

from pokerrl_b200.game.Poker import Poker
from pokerrl_b200.game.batched_env import BatchedPokerEnv

# field order of csrc/env_kernels.cu (enum Field)
_F = {k: i for i, k in enumerate(
    ["round", "pot", "stack0", "stack1", "bet0", "bet1", "flags", "cur", "last_raiser", "n_act_ep", "n_raises", "capped",
     "cap_raiser", "cap_noreopen", "last_type", "last_amt", "last_who", "done"])}
_FL_ALLIN, _FL_FOLD, _FL_ACTED = (1, 2), (4, 8), (16, 32)


class PokerEnv:
    def __init__(self, game_cls, env_args, lut_holder=None, is_evaluating=True, device=None, seed=0):
        if not is_evaluating:
            raise NotImplementedError("training-mode stack randomisation / action interpolation is not part of the tabular path")
        self._game, self.env_args = game_cls, env_args
        self.rules = game_cls.RULES
        self.lut_holder = lut_holder if lut_holder is not None else self.rules.get_lut_holder()
        self._b = BatchedPokerEnv(game_cls, env_args, 1, device=device, seed=seed)
        self.N_SEATS, self.N_ACTIONS = 2, self._b.N_ACTIONS
        self.N_CARDS_IN_DECK = self.rules.N_CARDS_IN_DECK
        self.IS_EVALUATING = True
        self._last = None
        self._host = None  # host copy of (state, deck) of the table, refreshed lazily after reset / step / load_state_dict
        # constants the evaluators read from the env (PokerEnv.py:361-368, games.py)
        self.REWARD_SCALAR = float(self._b.cfg.reward_scalar)
        self.EV_NORMALIZER = game_cls.EV_NORMALIZER
        self.IS_FIXED_LIMIT_GAME = bool(game_cls.IS_FIXED_LIMIT_GAME)
        self.bet_sizes_list_as_frac_of_pot = (sorted(env_args.bet_sizes_list_as_frac_of_pot)
                                              if game_cls.BETTING == "discretized" else None)

    # ---- PokerEnv.reset (:1075-1122)
    def reset(self, deck_state_dict=None):
        decks = None
        if deck_state_dict is not None:
            decks = self._deck_from_cards_state_dict(deck_state_dict)[None]
        obs, legal = self._b.reset(decks=decks)
        self._host = None
        self._last = (obs[0].cpu().numpy().copy(), np.zeros(2, np.float64), False)
        return self._last[0], np.zeros(self.N_SEATS, np.float32), False, [False, None]

    # ---- PokerEnv.step (:1148-1159)
    def step(self, action):
        obs, rew, done, legal = self._b.step(torch.tensor([int(action)], dtype=torch.int32))
        o, r, d = obs[0].cpu().numpy().copy(), rew[0].cpu().numpy().copy(), bool(done[0].item())
        self._host = None
        self._last = (o, r, d)
        return o, r, d, [False, None]

    def step_raise_pot_frac(self, pot_frac):
        """PokerEnv.step_raise_pot_frac (:1124-1136) for the fractions of this table's own bet set (fixed-limit games: the one
        raise there is); other fractions would need the continuous action space, which the device engine does not have"""
        if self.bet_sizes_list_as_frac_of_pot is None:
            return self.step(Poker.BET_RAISE)
        for i, f in enumerate(self.bet_sizes_list_as_frac_of_pot):
            if abs(float(f) - float(pot_frac)) <= 1e-9 * max(1.0, abs(float(f))):
                return self.step(2 + i)
        raise ValueError("pot fraction %r is not in this table's bet set %r" % (pot_frac, self.bet_sizes_list_as_frac_of_pot))

    # ---- read-only views of the table the evaluators use (PokerEnv attributes / PokerPlayer fields)
    def _state(self):
        if self._host is None:
            self._host = (self._b.state[:, 0].cpu().numpy(), self._b.deck[0].cpu().numpy())
        return self._host

    class _Seat:
        def __init__(self, seat_id, stack, current_bet):
            self.seat_id, self.stack, self.current_bet = seat_id, stack, current_bet

    @property
    def current_round(self):
        return int(self._state()[0][_F["round"]])

    @property
    def current_player(self):
        st = self._state()[0]
        p = int(st[_F["cur"]])
        return PokerEnv._Seat(p, int(st[_F["stack0"] + p]), int(st[_F["bet0"] + p]))

    @property
    def seats(self):
        st = self._state()[0]
        return [PokerEnv._Seat(p, int(st[_F["stack0"] + p]), int(st[_F["bet0"] + p])) for p in range(2)]

    @property
    def board(self):
        """int8 [N_TOTAL_BOARD_CARDS, 2] (rank, suit), not-dealt tokens where nothing lies yet (PokerEnv.board)"""
        st, d = self._state()
        nh, nbc = self._deal_layout()
        b = np.full(nbc, Poker.CARD_NOT_DEALT_TOKEN_1D, np.int8)
        n_out = self.rules.n_cards_out_at(int(st[_F["round"]]))
        b[:n_out] = d[2 * nh:2 * nh + n_out]
        return self.lut_holder.get_2d_cards(b)

    def get_all_winnable_money(self):  # main pot + side pots + the bets in front of the players (PokerEnv.py)
        st = self._state()[0]
        return int(st[_F["pot"]]) + int(st[_F["bet0"]]) + int(st[_F["bet1"]])

    def get_hole_cards_of_player(self, p_id):
        nh, _ = self._deal_layout()
        return self.lut_holder.get_2d_cards(self._state()[1][p_id * nh:(p_id + 1) * nh])

    def get_range_idx(self, p_id):
        return int(self.lut_holder.get_range_idx_from_hole_cards(self.get_hole_cards_of_player(p_id)))

    def get_legal_actions(self):  # DiscretizedPokerEnv.get_legal_actions (:99-135) / LimitPokerEnv (:41-59)
        return [int(a) for a in np.nonzero(self._b.legal[0].cpu().numpy())[0]]

    def get_current_obs(self, is_terminal=False):  # PokerEnv.get_current_obs (:1253-1271)
        return np.zeros_like(self._last[0]) if is_terminal else self._last[0]

    # ---- cards: the device deck holds the whole deal, top card first: seat 0's hole cards, seat 1's, then the board
    def _deal_layout(self):
        nh, r = self.rules.N_HOLE_CARDS, self.rules
        return nh, r.N_TOTAL_BOARD_CARDS

    def _deck_from_cards_state_dict(self, csd):
        """PokerEnv.load_cards_state_dict (:1367-1371): hands + the not yet dealt rest of the deck (drawn from the top)"""
        lut = self.lut_holder
        hands = [np.asarray(lut.get_1d_cards(np.asarray(h))).reshape(-1) for h in csd["hand"]]
        rest = np.asarray(lut.get_1d_cards(np.asarray(csd["deck"]["deck_remaining"]))).reshape(-1)
        deck = np.concatenate(hands + [rest]).astype(np.int8)
        assert deck.size == self.N_CARDS_IN_DECK and len(set(deck.tolist())) == deck.size
        return deck

    def cards_state_dict(self):  # :1361-1365 (right after reset: nothing dealt to the board yet)
        nh, _ = self._deal_layout()
        d = self._b.deck[0].cpu().numpy()
        lut = self.lut_holder
        return {"deck": {"deck_remaining": lut.get_2d_cards(d[2 * nh:])},
                "board": np.full((self.rules.N_TOTAL_BOARD_CARDS, 2), Poker.CARD_NOT_DEALT_TOKEN_1D, np.int8),
                "hand": [lut.get_2d_cards(d[p * nh:(p + 1) * nh]) for p in range(2)]}

    # ---- PokerEnv.state_dict / load_state_dict (:1161-1251)
    def state_dict(self):
        st = self._b.state[:, 0].cpu().numpy()
        d = self._b.deck[0].cpu().numpy()
        nh, nbc = self._deal_layout()
        rnd = int(st[_F["round"]])
        n_out = self.rules.n_cards_out_at(rnd)
        board = np.full(nbc, Poker.CARD_NOT_DEALT_TOKEN_1D, np.int8)
        board[:n_out] = d[2 * nh:2 * nh + n_out]
        flags = int(st[_F["flags"]])
        lut = self.lut_holder
        lt = int(st[_F["last_type"]])
        return {
            "current_round": rnd, "side_pots": [0, 0], "main_pot": int(st[_F["pot"]]),
            "board_2d": lut.get_2d_cards(board), "board_1d": board,
            "last_action": [None, None, None] if lt < 0 else [lt, int(st[_F["last_amt"]]), int(st[_F["last_who"]])],
            "capped_raise": {"happened_this_round": bool(st[_F["capped"]]), "player_that_raised": int(st[_F["cap_raiser"]]),
                             "player_that_cant_reopen": int(st[_F["cap_noreopen"]])},
            "current_player": int(st[_F["cur"]]),
            "last_raiser": None if st[_F["last_raiser"]] < 0 else int(st[_F["last_raiser"]]),
            "deck_remaining": d[2 * nh + n_out:].copy(), "n_actions_this_episode": int(st[_F["n_act_ep"]]),
            "n_raises_this_round": int(st[_F["n_raises"]]), "is_evaluating": True,
            "seats": [{"seat_id": p, "hand": lut.get_2d_cards(d[p * nh:(p + 1) * nh]), "stack": int(st[_F["stack0"] + p]),
                       "current_bet": int(st[_F["bet0"] + p]), "is_allin": bool(flags & _FL_ALLIN[p]),
                       "folded_this_episode": bool(flags & _FL_FOLD[p]), "has_acted_this_round": bool(flags & _FL_ACTED[p]),
                       "side_pot_rank": -1} for p in range(2)],
            "_raw": {"state": st.copy(), "deck": d.copy(), "legal": self._b.legal[0].cpu().numpy().copy(),
                     "last": None if self._last is None else tuple(np.copy(x) if isinstance(x, np.ndarray) else x for x in self._last)},
        }

    def load_state_dict(self, env_state_dict, blank_private_info=False):
        raw = env_state_dict["_raw"]
        self._b.state[:, 0] = torch.from_numpy(raw["state"]).to(self._b.device)
        self._b.deck[0] = torch.from_numpy(raw["deck"]).to(self._b.device)
        self._b.legal[0] = torch.from_numpy(raw["legal"]).to(self._b.device)
        self._host = None
        self._last = raw["last"]
