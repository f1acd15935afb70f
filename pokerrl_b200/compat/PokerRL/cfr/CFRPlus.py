from pokerrl_b200.cfr.CFRPlus import CFRPlus  # noqa: F401
