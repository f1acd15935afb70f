from pokerrl_b200.game.look_up_table import LutHolder, LutHolderHoldem, LutHolderLeduc  # noqa: F401
