from pokerrl_b200.eval.br.LocalBRMaster import LocalBRMaster  # noqa: F401
