from pokerrl_b200.eval.lbr.LocalLBRMaster import LocalLBRMaster  # noqa: F401
