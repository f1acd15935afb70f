from pokerrl_b200.game.PublicTree import PublicTree  # noqa: F401
