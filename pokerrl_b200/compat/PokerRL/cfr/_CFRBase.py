from pokerrl_b200.cfr._CFRBase import CFRBase  # noqa: F401
