from pokerrl_b200.rl.base_cls.EvalAgentBase import EvalAgentBase  # noqa: F401
