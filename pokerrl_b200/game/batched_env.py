"""Batched heads-up PokerEnv on the GPU: `reset()` / `step(actions)` for n_envs tables at once (host side of
csrc/env_kernels.cu; replaces the scalar `PokerRL/game/_/rl_env/base/PokerEnv.py` reset/step loop).

Return convention per call mirrors `PokerEnv.step` (PokerEnv.py:1148-1159) batched: obs float32 [B, obs_size]
(zeros at terminal states), rewards float64 [B, 2], done uint8 [B], plus the legal-action mask uint8 [B, N_ACTIONS] of
the NEXT decision (`get_legal_actions`)."""
import ctypes as C

import torch

from pokerrl_b200 import _native as nat
from pokerrl_b200.game.Poker import Poker


def env_config(game_cls, env_args, n_envs):
    g, r = game_cls, game_cls.RULES
    c = nat.PrlEnvCfg()
    c.n_envs = n_envs
    if g.BETTING == "limit":
        c.kind = 0
    elif g.BETTING == "discretized":
        c.kind = 1
    else:
        raise NotImplementedError("continuous no-limit action spaces take (type, chips) tuples; use a discretized game")
    c.n_actions = env_args.N_ACTIONS
    c.n_rounds, c.n_round_slots = len(r.ALL_ROUNDS_LIST), r.ALL_ROUNDS_LIST[-1] + 1
    c.n_hole, c.n_ranks, c.n_suits, c.n_deck = r.N_HOLE_CARDS, r.N_RANKS, r.N_SUITS, r.N_CARDS_IN_DECK
    c.n_flop, c.n_turn, c.n_river = r.N_FLOP_CARDS, r.N_TURN_CARDS, r.N_RIVER_CARDS
    c.small_blind, c.big_blind, c.ante = g.SMALL_BLIND, g.BIG_BLIND, g.ANTE
    c.small_bet, c.big_bet = g.SMALL_BET or 0, g.BIG_BET or 0
    c.round_big_bet_starts = g.ROUND_WHERE_BIG_BET_STARTS if g.ROUND_WHERE_BIG_BET_STARTS is not None else 0
    for rnd in range(4):
        c.max_raises[rnd] = (g.MAX_N_RAISES_PER_ROUND or {}).get(rnd, 0)
    c.first_action_no_call, c.limit_raise_is_pot = int(g.FIRST_ACTION_NO_CALL), int(g.LIMIT_RAISE_IS_POT)
    c.btn_first_postflop, c.suits_matter = int(r.BTN_IS_FIRST_POSTFLOP), int(r.SUITS_MATTER)
    c.pair_bonus = r.PAIR_BONUS or 0
    stacks = [g.DEFAULT_STACK_SIZE if s is None else int(s) for s in env_args.starting_stack_sizes_list]
    c.start_stack[0], c.start_stack[1] = stacks
    c.obs_size = 7 + 3 + 2 + 2 + c.n_round_slots + 6 + r.N_TOTAL_BOARD_CARDS * (r.N_RANKS + r.N_SUITS)
    if c.kind == 1:
        for i, f in enumerate(sorted(env_args.bet_sizes_list_as_frac_of_pot)):  # DiscretizedPokerEnv.py:40
            c.fracs[i] = float(f)
    mean_stack = float(sum(stacks)) / 2
    c.reward_scalar = mean_stack / 5 if env_args.scale_rewards else 1.0  # PokerEnv.py:361-368
    c.norm = mean_stack  # PokerEnv.py:1267
    return c


class BatchedPokerEnv:
    def __init__(self, game_cls, env_args, n_envs, device=None, seed=0):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedPokerEnv needs a CUDA device; there is no CPU fallback")
        assert env_args.n_seats == 2
        self.cfg = env_config(game_cls, env_args, n_envs)
        device = device if device is not None else "cuda:%d" % torch.cuda.current_device()
        self.n_envs, self.device, self.seed = n_envs, torch.device(device), int(seed)
        nf = nat.lib().prl_env_state_fields()
        z = lambda *s, dtype: torch.zeros(*s, dtype=dtype, device=self.device)  # noqa: E731
        self.state = z(nf, n_envs, dtype=torch.int32)
        self.deck = z(n_envs, self.cfg.n_deck, dtype=torch.int8)
        self.obs = z(n_envs, self.cfg.obs_size, dtype=torch.float32)
        self.rewards = z(n_envs, 2, dtype=torch.float64)
        self.done = z(n_envs, dtype=torch.uint8)
        self.legal = z(n_envs, self.cfg.n_actions, dtype=torch.uint8)
        self._step_id, self._episode0 = 0, 0
        self.N_ACTIONS, self.obs_size = self.cfg.n_actions, self.cfg.obs_size

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, decks=None):
        """decks: optional int8 [B, n_deck] (top card first: seat 0's hole cards, seat 1's, then the board) - the
        batched analogue of `reset(deck_state_dict=...)` (PokerEnv.py:1118-1120); otherwise shuffled on the device."""
        if decks is not None:
            self.deck.copy_(torch.as_tensor(decks).to(device=self.device, dtype=torch.int8))
        with torch.cuda.device(self.device):
            nat.call("prl_env_reset", C.byref(self.cfg), C.c_void_p(self.state.data_ptr()), C.c_void_p(self.deck.data_ptr()),
                     C.c_void_p(self.obs.data_ptr()), C.c_void_p(self.legal.data_ptr()), self.seed, self._episode0,
                     int(decks is None), self._stream())
        self._episode0 += self.n_envs
        return self.obs, self.legal

    def step(self, actions=None, auto_reset=False):
        """actions: int32 [B] discrete actions (0 fold, 1 check/call, 2.. raises); None / negative = uniformly random
        legal action.  Tables that are done ignore the step unless auto_reset."""
        a_ptr = None
        if actions is not None:
            a = torch.as_tensor(actions).to(device=self.device, dtype=torch.int32).contiguous()
            a_ptr = C.c_void_p(a.data_ptr())
        with torch.cuda.device(self.device):
            nat.call("prl_env_step", C.byref(self.cfg), C.c_void_p(self.state.data_ptr()), C.c_void_p(self.deck.data_ptr()),
                     a_ptr, C.c_void_p(self.obs.data_ptr()), C.c_void_p(self.rewards.data_ptr()),
                     C.c_void_p(self.done.data_ptr()), C.c_void_p(self.legal.data_ptr()), self.seed, self._step_id,
                     int(auto_reset), self._stream())
        self._step_id += 1
        return self.obs, self.rewards, self.done, self.legal
