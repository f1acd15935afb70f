from pokerrl_b200.game.wrappers import *  # noqa: F401,F403
