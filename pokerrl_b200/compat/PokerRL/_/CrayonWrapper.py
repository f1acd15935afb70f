from pokerrl_b200._.CrayonWrapper import CrayonWrapper  # noqa: F401
