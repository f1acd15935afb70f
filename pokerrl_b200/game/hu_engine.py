"""Heads-up betting state machine (host side, integer chips, no cards).

Restates the *public* part of the reference's `PokerEnv` for two seats:
  reset            PokerEnv.py:1075-1122        legalisation     PokerEnv.py:885-941
  step / apply     PokerEnv.py:681-732          round end test   PokerEnv.py:943-954
  transitions      PokerEnv.py:737-789          min raise        PokerEnv.py:809-812
  pot bookkeeping  PokerEnv.py:539-551          pot-fraction     PokerEnv.py:1376-1396
  discretized      DiscretizedPokerEnv.py:44-135    limit        LimitPokerEnv.py:27-59, games.py:253-254
Cards never influence betting legality, so the tree compiler (`flat_tree.py`) enumerates this machine once and
expands the result over boards analytically, and the batched CUDA env (`csrc/env_kernels.cu`) implements the very
same transition function per lane.
"""
from pokerrl_b200.game.Poker import Poker

FOLD, CALL, RAISE = Poker.FOLD, Poker.CHECK_CALL, Poker.BET_RAISE

# outcome codes of HUBetting.step
CONTINUE, NEXT_ROUND, SHOWDOWN, ALLIN_RUNDOWN, FOLDED = 0, 1, 2, 3, 4


class HUState:
    __slots__ = ("round", "main_pot", "stack", "bet", "allin", "folded", "acted", "cur", "last_raiser",
                 "n_actions_ep", "n_raises_round", "capped", "capped_raiser", "capped_cant_reopen", "last_action")

    def copy(self):
        s = HUState.__new__(HUState)
        s.round, s.main_pot, s.cur = self.round, self.main_pot, self.cur
        s.stack, s.bet = self.stack[:], self.bet[:]
        s.allin, s.folded, s.acted = self.allin[:], self.folded[:], self.acted[:]
        s.last_raiser, s.n_actions_ep, s.n_raises_round = self.last_raiser, self.n_actions_ep, self.n_raises_round
        s.capped, s.capped_raiser, s.capped_cant_reopen = self.capped, self.capped_raiser, self.capped_cant_reopen
        s.last_action = self.last_action
        return s


class HUBetting:
    def __init__(self, game_cls, env_args):
        assert env_args.n_seats == 2, "heads-up only (the tabular CFR/BR path of the reference is HU: ValueFiller.py:27)"
        g = game_cls
        self.game_cls = g
        self.kind = g.BETTING
        self.is_limit = g.IS_FIXED_LIMIT_GAME
        self.SB, self.BB, self.ANTE = g.SMALL_BLIND, g.BIG_BLIND, g.ANTE
        self.last_round = g.RULES.ALL_ROUNDS_LIST[-1]
        self.btn_first_postflop = g.RULES.BTN_IS_FIRST_POSTFLOP
        self.first_action_no_call = g.FIRST_ACTION_NO_CALL
        self.stacks0 = [g.DEFAULT_STACK_SIZE if s is None else int(s) for s in env_args.starting_stack_sizes_list]
        self.N_ACTIONS = env_args.N_ACTIONS
        if self.kind == "discretized":
            self.fracs = sorted(env_args.bet_sizes_list_as_frac_of_pot)  # DiscretizedPokerEnv.py:40

    # ------------------------------------------------------------------ reset
    def reset(self, stacks=None):
        s = HUState.__new__(HUState)
        s.n_raises_round = (1 if self.BB > 0 else 0) if self.is_limit else 0
        s.main_pot, s.round = 0, Poker.PREFLOP
        s.capped, s.capped_raiser, s.capped_cant_reopen = False, -1, -1
        s.last_raiser, s.n_actions_ep = -1, 0
        s.last_action = None
        s.stack = list(self.stacks0 if stacks is None else stacks)
        s.bet, s.allin, s.folded, s.acted = [0, 0], [False, False], [False, False], [False, False]
        # antes go straight to the pot (PokerEnv.py:1111-1112)
        for p in (0, 1):
            self._bet_to(s, p, self.ANTE)
            s.acted[p] = False
        self._bets_into_pot(s)
        self._bet_to(s, 0, self.SB)  # HU: seat 0 = BTN = SB, seat 1 = BB (PokerEnv.py:337-340)
        s.acted[0] = False
        self._bet_to(s, 1, self.BB)
        s.acted[1] = False
        s.cur = 0
        return s

    # ------------------------------------------------------------------ primitives
    @staticmethod
    def _bet_to(s, p, total):
        s.acted[p] = True
        s.stack[p] -= total - s.bet[p]
        s.bet[p] = total
        if s.stack[p] == 0:
            s.allin[p] = True

    @staticmethod
    def _bets_into_pot(s):
        d = s.bet[0] - s.bet[1]
        if d > 0:
            s.stack[0] += d
            s.bet[0] -= d
        elif d < 0:
            s.stack[1] -= d
            s.bet[1] += d
        s.main_pot += s.bet[0] + s.bet[1]
        s.bet[0] = s.bet[1] = 0

    def min_raise_total(self, s):
        lo, hi = (s.bet[0], s.bet[1]) if s.bet[0] <= s.bet[1] else (s.bet[1], s.bet[0])
        return hi + max(hi - lo, self.BB)

    @staticmethod
    def pot_fraction_raise(s, frac, p):
        to_call = max(s.bet) - s.bet[p]
        pot_after_call = s.main_pot + s.bet[0] + s.bet[1] + to_call
        return int(to_call + pot_after_call * frac) + s.bet[p]

    # ------------------------------------------------------------------ action decoding / legalisation
    def decode(self, s, a):
        """discrete action -> (type, chips) in PokerEnv's continuous form"""
        if a == FOLD:
            return FOLD, -1
        if a == CALL:
            return CALL, -1
        if self.kind == "discretized":
            return RAISE, self.pot_fraction_raise(s, self.fracs[a - 2], s.cur)
        if self.kind == "limit":
            if a != RAISE:
                raise ValueError(a)
            return RAISE, -1
        raise ValueError("no-limit envs take (type, chips) tuples; use step_tuple")

    def _adjust_raise(self, s, chips):
        if self.kind == "limit":
            if self.game_cls.LIMIT_RAISE_IS_POT:
                return self.pot_fraction_raise(s, 1.0, s.cur)
            g = self.game_cls
            b = g.BIG_BET if s.round >= g.ROUND_WHERE_BIG_BET_STARTS else g.SMALL_BET
            return (s.n_raises_round + 1) * b
        return max(self.min_raise_total(s), chips)

    def _check_call(self, s, total_to_call):
        p = s.cur
        return CALL, int(min(total_to_call - s.bet[p], s.stack[p]) + s.bet[p])

    def fix(self, s, typ, chips):
        """PokerEnv._get_fixed_action"""
        p = s.cur
        total_to_call = max(s.bet)
        if typ == FOLD:
            if total_to_call <= s.bet[p]:
                return self._check_call(s, total_to_call)
            return FOLD, -1
        if typ == CALL:
            if self.first_action_no_call and s.n_actions_ep == 0 and s.round == Poker.PREFLOP:
                return FOLD, -1
            return self._check_call(s, total_to_call)
        if typ == RAISE:
            if self.is_limit and s.n_raises_round >= self.game_cls.MAX_N_RAISES_PER_ROUND[s.round]:
                return self._check_call(s, total_to_call)
            if s.stack[p] + s.bet[p] <= total_to_call or (s.capped and s.capped_cant_reopen == p):
                return self._check_call(s, total_to_call)
            raise_to = self._adjust_raise(s, chips)
            if s.bet[p] + s.stack[p] < raise_to:
                raise_to = s.stack[p] + s.bet[p]
            return RAISE, int(raise_to)
        raise RuntimeError("invalid action type %r" % (typ,))

    def legal_actions(self, s):
        legal = []
        for a in (FOLD, CALL):
            if self.fix(s, a, -1)[0] == a:
                legal.append(a)
        if self.kind == "limit":
            if (s.n_raises_round < self.game_cls.MAX_N_RAISES_PER_ROUND[s.round]
                    and self.fix(s, RAISE, -1)[0] == RAISE):
                legal.append(RAISE)
            return legal
        if self.kind == "nolimit":
            if self.fix(s, RAISE, 1)[0] == RAISE:
                legal.append(RAISE)
            return legal
        # discretized: ascending raise sizes, collapse sizes that round up to the min-raise, stop after the
        # first size that is capped down to all-in (DiscretizedPokerEnv.py:99-135)
        last_too_small = None
        for a in range(2, self.N_ACTIONS):
            t, want = self.decode(s, a)
            ft, got = self.fix(s, t, want)
            if ft != t:
                break
            if want < got:
                last_too_small = a
            else:
                if last_too_small is not None:
                    legal.append(last_too_small)
                    last_too_small = None
                legal.append(a)
            if want > got:
                break
        assert legal
        return legal

    # ------------------------------------------------------------------ step
    def step(self, s, a):
        """Applies discrete action `a` in place. Returns (outcome, pre) where `pre` is a copy of the state
        'before money moves' for NEXT_ROUND (bets still in front), None for CONTINUE, and for terminal outcomes
        `s` itself holds the state after bets were pushed into the pot but before the payout."""
        typ, chips = self.decode(s, a)
        return self.step_tuple(s, typ, chips)

    def step_tuple(self, s, typ, chips):
        typ, amt = self.fix(s, typ, chips)
        p = s.cur
        if typ == CALL:
            self._bet_to(s, p, amt)
        elif typ == FOLD:
            s.acted[p] = True
            s.folded[p] = True
        else:
            if amt < self.min_raise_total(s):  # under-min all-in: previous raiser may not re-open
                s.capped, s.capped_raiser, s.capped_cant_reopen = True, p, s.last_raiser
            elif s.capped and s.capped_cant_reopen != p:
                s.capped, s.capped_raiser, s.capped_cant_reopen = False, -1, -1
            s.last_raiser = p
            self._bet_to(s, p, amt)
            s.n_actions_ep += 1
            if self.is_limit:
                s.n_raises_round += 1
        s.last_action = (typ, amt, p)

        n_nonfold = (not s.folded[0]) + (not s.folded[1])
        live = [q for q in (0, 1) if not s.folded[q] and not s.allin[q]]
        if self._continue_round(s, n_nonfold, live):
            q = 1 - p
            s.cur = q if (not s.allin[q] and not s.folded[q]) else p
            return CONTINUE, None
        if len(live) > 1:
            if s.round == self.last_round:
                self._bets_into_pot(s)
                return SHOWDOWN, None
            pre = s.copy()
            self._next_round(s)
            return NEXT_ROUND, pre
        if n_nonfold > 1:
            self._bets_into_pot(s)
            return ALLIN_RUNDOWN, None
        self._bets_into_pot(s)
        return FOLDED, None

    @staticmethod
    def _continue_round(s, n_nonfold, live):
        if n_nonfold < 2:
            return False
        largest = max(s.bet)
        settled = all(s.folded[q] or s.allin[q] or s.bet[q] == largest for q in (0, 1))
        return not (settled and all(s.acted[q] for q in live))

    def _next_round(self, s):
        if self.is_limit:
            s.n_raises_round = 0
        s.capped, s.capped_raiser, s.capped_cant_reopen = False, -1, -1
        self._bets_into_pot(s)
        s.cur = 0 if self.btn_first_postflop else 1
        s.acted = [False, False]
        s.round += 1
