from pokerrl_b200.eval.lbr.LocalLBRWorker import LocalLBRWorker  # noqa: F401
