from pokerrl_b200.game.Poker import Poker  # noqa: F401
