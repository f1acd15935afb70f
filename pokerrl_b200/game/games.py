"""Game definitions: rule sets (deck / hand / board geometry) and betting constants.

Declarative mirror of the reference's rule classes (`PokerRL/game/_/rl_env/game_rules.py:15-313`) and game
classes (`PokerRL/game/games.py:18-269`).  A game class here is *data only* – the betting semantics live in
`hu_engine.HUBetting` and the value/regret arithmetic lives in the CUDA library.  Attribute names are the
reference's, so `game_cls.DEFAULT_STACK_SIZE`, `game_cls.EV_NORMALIZER`, `game_cls.ARGS_CLS(...)`,
`game_cls.RULES.RANGE_SIZE`, … keep working for code written against PokerRL.
"""
from math import comb

from pokerrl_b200.game.Poker import Poker
from pokerrl_b200.game.poker_env_args import DiscretizedPokerEnvArgs, LimitPokerEnvArgs, NoLimitPokerEnvArgs


# ------------------------------------------------------------------------------------------------ rule sets
class _Rules:
    N_HOLE_CARDS = N_RANKS = N_SUITS = None
    N_FLOP_CARDS = N_TURN_CARDS = N_RIVER_CARDS = 0
    ALL_ROUNDS_LIST = [Poker.PREFLOP, Poker.FLOP]
    BTN_IS_FIRST_POSTFLOP = False
    SUITS_MATTER = True
    STRING = None
    # value added to the rank of a hand that pairs the board in the one-card games (game_rules.py:68-75,134-141)
    PAIR_BONUS = None

    @classmethod
    def n_cards_dealt_in_transition_to(cls, rnd):
        return {Poker.PREFLOP: 0, Poker.FLOP: cls.N_FLOP_CARDS, Poker.TURN: cls.N_TURN_CARDS,
                Poker.RIVER: cls.N_RIVER_CARDS}[rnd]

    @classmethod
    def n_cards_out_at(cls, rnd):
        return sum(cls.n_cards_dealt_in_transition_to(r) for r in range(rnd + 1))

    @classmethod
    def get_lut_holder(cls):
        from pokerrl_b200.game.look_up_table import LutHolder
        return LutHolder(cls)


def _finish(rules):
    rules.N_CARDS_IN_DECK = rules.N_RANKS * rules.N_SUITS
    rules.RANGE_SIZE = comb(rules.N_CARDS_IN_DECK, rules.N_HOLE_CARDS)
    rules.N_TOTAL_BOARD_CARDS = rules.N_FLOP_CARDS + rules.N_TURN_CARDS + rules.N_RIVER_CARDS
    rules.ROUND_BEFORE = {r: max(r - 1, 0) for r in rules.ALL_ROUNDS_LIST}
    rules.ROUND_AFTER = {r: (r + 1 if r + 1 in rules.ALL_ROUNDS_LIST else None) for r in rules.ALL_ROUNDS_LIST}
    return rules


@_finish
class LeducRules(_Rules):
    """3 ranks x 2 suits, one hole card, one board card (game_rules.py:15-81)."""
    N_HOLE_CARDS, N_RANKS, N_SUITS = 1, 3, 2
    N_FLOP_CARDS = 1
    BTN_IS_FIRST_POSTFLOP = True
    SUITS_MATTER = False
    STRING = "LEDUC_RULES"
    PAIR_BONUS = 100


@_finish
class BigLeducRules(_Rules):
    """12 ranks x 2 suits variant (game_rules.py:84-147)."""
    N_HOLE_CARDS, N_RANKS, N_SUITS = 1, 12, 2
    N_FLOP_CARDS = 1
    BTN_IS_FIRST_POSTFLOP = True
    SUITS_MATTER = False
    STRING = "BIG_LEDUC_RULES"
    PAIR_BONUS = 10000


@_finish
class HoldemRules(_Rules):
    """Texas Hold'em: 52 cards, two hole cards, 3+1+1 board (game_rules.py:150-229)."""
    N_HOLE_CARDS, N_RANKS, N_SUITS = 2, 13, 4
    N_FLOP_CARDS, N_TURN_CARDS, N_RIVER_CARDS = 3, 1, 1
    ALL_ROUNDS_LIST = [Poker.PREFLOP, Poker.FLOP, Poker.TURN, Poker.RIVER]
    STRING = "HOLDEM_RULES"


@_finish
class FlopHoldemRules(_Rules):
    """Flop Hold'em ("FHP"): all five board cards are dealt at once after the pre-flop round
    (game_rules.py:232-313)."""
    N_HOLE_CARDS, N_RANKS, N_SUITS = 2, 13, 4
    N_FLOP_CARDS = 5
    STRING = "FLOP_HOLDEM_RULES"


# ------------------------------------------------------------------------------------------------ games
class _Game:
    RULES = None
    BETTING = None  # "limit" | "discretized" | "nolimit"
    ARGS_CLS = None
    IS_FIXED_LIMIT_GAME = False
    IS_POT_LIMIT_GAME = False
    FIRST_ACTION_NO_CALL = False  # PokerEnv.py:87
    LIMIT_RAISE_IS_POT = False  # Flop5Holdem overrides the limit raise with a pot-size raise (games.py:253-254)
    SMALL_BLIND = BIG_BLIND = ANTE = 0
    SMALL_BET = BIG_BET = None
    MAX_N_RAISES_PER_ROUND = None
    ROUND_WHERE_BIG_BET_STARTS = None
    DEFAULT_STACK_SIZE = None
    EV_NORMALIZER = None
    WIN_METRIC = None

    @classmethod
    def get_lut_holder(cls):
        return cls.RULES.get_lut_holder()


class StandardLeduc(_Game):
    """games.py:18-48"""
    RULES, BETTING, ARGS_CLS = LeducRules, "limit", LimitPokerEnvArgs
    IS_FIXED_LIMIT_GAME = True
    MAX_N_RAISES_PER_ROUND = {Poker.PREFLOP: 2, Poker.FLOP: 2}
    ANTE, SMALL_BET, BIG_BET = 1, 2, 4
    DEFAULT_STACK_SIZE = 13
    EV_NORMALIZER = 1000.0 / ANTE
    WIN_METRIC = Poker.MeasureAnte
    ROUND_WHERE_BIG_BET_STARTS = Poker.FLOP


class BigLeduc(_Game):
    """games.py:51-77"""
    RULES, BETTING, ARGS_CLS = BigLeducRules, "limit", LimitPokerEnvArgs
    IS_FIXED_LIMIT_GAME = True
    MAX_N_RAISES_PER_ROUND = {Poker.PREFLOP: 6, Poker.FLOP: 6}
    ANTE, SMALL_BET, BIG_BET = 1, 2, 4
    DEFAULT_STACK_SIZE = 100
    EV_NORMALIZER = 1000.0 / ANTE
    WIN_METRIC = Poker.MeasureAnte
    ROUND_WHERE_BIG_BET_STARTS = Poker.FLOP


class _BlindGame(_Game):
    SMALL_BLIND, BIG_BLIND = 50, 100
    DEFAULT_STACK_SIZE = 20000
    EV_NORMALIZER = 1000.0 / BIG_BLIND
    WIN_METRIC = Poker.MeasureBB


class NoLimitLeduc(_BlindGame):
    """games.py:80-103"""
    RULES, BETTING, ARGS_CLS = LeducRules, "nolimit", NoLimitPokerEnvArgs


class DiscretizedNLLeduc(_BlindGame):
    """games.py:106-129"""
    RULES, BETTING, ARGS_CLS = LeducRules, "discretized", DiscretizedPokerEnvArgs


class LimitHoldem(_Game):
    """games.py:134-167"""
    RULES, BETTING, ARGS_CLS = HoldemRules, "limit", LimitPokerEnvArgs
    IS_FIXED_LIMIT_GAME = True
    MAX_N_RAISES_PER_ROUND = {Poker.PREFLOP: 4, Poker.FLOP: 4, Poker.TURN: 4, Poker.RIVER: 4}
    ROUND_WHERE_BIG_BET_STARTS = Poker.TURN
    SMALL_BLIND, BIG_BLIND, SMALL_BET, BIG_BET = 1, 2, 2, 4
    DEFAULT_STACK_SIZE = 48
    EV_NORMALIZER = 1000.0 / BIG_BLIND
    WIN_METRIC = Poker.MeasureBB


class NoLimitHoldem(_BlindGame):
    """games.py:170-195"""
    RULES, BETTING, ARGS_CLS = HoldemRules, "nolimit", NoLimitPokerEnvArgs


class DiscretizedNLHoldem(_BlindGame):
    """games.py:198-219"""
    RULES, BETTING, ARGS_CLS = HoldemRules, "discretized", DiscretizedPokerEnvArgs


class Flop5Holdem(_BlindGame):
    """games.py:222-254: limit-type action space {fold, call, raise} whose raise is pot-sized, one raise per
    round (the big blind counts as the first), SB may not limp."""
    RULES, BETTING, ARGS_CLS = FlopHoldemRules, "limit", LimitPokerEnvArgs
    IS_FIXED_LIMIT_GAME = True
    MAX_N_RAISES_PER_ROUND = {Poker.PREFLOP: 2, Poker.FLOP: 2}
    ROUND_WHERE_BIG_BET_STARTS = Poker.TURN
    FIRST_ACTION_NO_CALL = True
    LIMIT_RAISE_IS_POT = True


ALL_ENVS = [StandardLeduc, BigLeduc, NoLimitLeduc, DiscretizedNLLeduc, LimitHoldem, NoLimitHoldem,
            DiscretizedNLHoldem, Flop5Holdem]


def get_env_cls_from_str(env_str):
    """rl_util.py:75-80"""
    for e in ALL_ENVS:
        if e.__name__ == env_str:
            return e
    raise ValueError(env_str, "is not registered or does not exist.")
