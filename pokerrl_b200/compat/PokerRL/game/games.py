from pokerrl_b200.game.games import *  # noqa: F401,F403
