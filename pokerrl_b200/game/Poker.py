"""Global poker constants. Same names and values as the reference's `PokerRL/game/Poker.py:6-46`
so that actions, round ids, the not-dealt token and metric names are interchangeable."""


class Poker:
    PREFLOP, FLOP, TURN, RIVER = 0, 1, 2, 3
    INT2STRING_ROUND = {0: "preflop", 1: "flop", 2: "turn", 3: "river"}
    STRING2INT_ROUND = {v: k for k, v in INT2STRING_ROUND.items()}

    FOLD, CHECK_CALL, BET_RAISE = 0, 1, 2

    CARD_NOT_DEALT_TOKEN_1D = -127

    MeasureAnte = "MA_per_G"  # milli-antes per game
    MeasureBB = "MBB_per_G"  # milli-big-blinds per game
