from pokerrl_b200.cfr.LinearCFR import LinearCFR  # noqa: F401
