from pokerrl_b200.cfr.VanillaCFR import VanillaCFR  # noqa: F401
