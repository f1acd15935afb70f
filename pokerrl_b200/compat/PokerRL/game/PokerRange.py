from pokerrl_b200.game.PokerRange import PokerRange  # noqa: F401
