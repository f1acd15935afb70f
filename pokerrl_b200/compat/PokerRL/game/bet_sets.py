from pokerrl_b200.game.bet_sets import *  # noqa: F401,F403
