from pokerrl_b200.eval.lbr.LBRArgs import LBRArgs  # noqa: F401
