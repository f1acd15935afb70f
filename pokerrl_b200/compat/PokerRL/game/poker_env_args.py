from pokerrl_b200.game.poker_env_args import *  # noqa: F401,F403
