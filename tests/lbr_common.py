"""Helpers of the LBR episode replay test: the deterministic hand-dependent policy the fixture generator
(oracle/gen_golden_lbr_run.py) gave the REFERENCE's LocalLBRWorker, as an eval agent of this package, with the recorded
uniform random numbers fed back in."""
import numpy as np

from pokerrl_b200.rl.base_cls.EvalAgentBase import EvalAgentBase


def policy_table(range_size, n_actions, legal, street):
    """float32 [R, N_ACTIONS]: weight 1 + ((7 h + 13 a + 3 street) mod 5) on the legal actions, rows normalised"""
    h = np.arange(range_size, dtype=np.int64)[:, None]
    a = np.arange(n_actions, dtype=np.int64)[None, :]
    w = (1 + ((7 * h + 13 * a + 3 * street) % 5)).astype(np.float32)
    mask = np.zeros(n_actions, np.float32)
    mask[list(legal)] = 1.0
    w = w * mask[None, :]
    return (w / w.sum(axis=1, keepdims=True)).astype(np.float32)


class ReplayTableAgent(EvalAgentBase):
    ALL_MODES = ["table"]

    def __init__(self, t_prof, mode=None, device=None):
        super().__init__(t_prof=t_prof, mode=mode, device=device)
        self.draws = []  # the uniform numbers the reference's agent consumed in the hand being replayed

    def can_compute_mode(self):
        return True

    def update_weights(self, weights_for_eval_agent):
        pass

    def _state_dict(self):
        return {}

    def _load_state_dict(self, state):
        pass

    def _uniform(self):
        return self.draws.pop(0)

    def get_a_probs_for_each_hand(self):
        env = self.internal_env
        return policy_table(self.env_bldr.rules.RANGE_SIZE, self.env_bldr.N_ACTIONS, env.get_legal_actions(), env.current_round)
