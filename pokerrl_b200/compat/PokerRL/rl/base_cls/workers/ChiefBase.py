from pokerrl_b200.rl.base_cls.workers.ChiefBase import ChiefBase  # noqa: F401
