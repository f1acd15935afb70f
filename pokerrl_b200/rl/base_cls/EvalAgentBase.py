"""Agent interface queried by the evaluators (`PokerRL/rl/base_cls/EvalAgentBase.py:9-170`).

Kept: modes, `get_a_probs_for_each_hand()` -> [RANGE_SIZE, N_ACTIONS], `set_to_public_tree_node_state(node)`,
`update_weights`, `can_compute_mode`, `state_dict` / `store_to_disk` / `load_from_disk`.  The reference replays the
observation history root->node through an env wrapper to feed neural nets (RecurrentHistoryWrapper.py:57-84); that
formatting is out of scope here, so the base class simply remembers the public-tree node it was set to."""
import pickle

from pokerrl_b200.rl.base_cls.TrainingProfileBase import get_env_builder


class EvalAgentBase:
    ALL_MODES = NotImplementedError

    def __init__(self, t_prof, mode=None, device=None):
        self.t_prof = t_prof
        self.env_bldr = get_env_builder(t_prof=t_prof)
        self._mode = mode
        self.device = device if device is not None else getattr(t_prof, "device_inference", None)
        self._node = None
        self._stack_size = None

    # ---- queries
    def get_a_probs_for_each_hand(self):
        raise NotImplementedError

    def get_a_probs_for_public_tree(self, tree):
        """Optional batched form of the query loop of StrategyFiller._fill_with_agent_policy (:88-116): action
        probabilities of EVERY decision node of `tree` at once, float32 [n_decision, RANGE_SIZE, N_ACTIONS] in the order of
        `tree.decision_nodes()` (a torch tensor, ideally already on the tree's device).  Return None (default) to be queried
        node by node through set_to_public_tree_node_state / get_a_probs_for_each_hand."""
        return None

    def can_compute_mode(self):
        raise NotImplementedError

    # ---- the agent's own table (EvalAgentBase.py:29, 128-158).  Evaluators that PLAY against the agent (LBR,
    #      eval/lbr/LocalLBRWorker.py) keep it in step with theirs through these notifications; the table is a single-table view of
    #      the device engine (game/poker_env.py), created on first use.
    @property
    def internal_env(self):
        if getattr(self, "_internal_env", None) is None:
            self._internal_env = self.env_bldr.get_new_env(is_evaluating=True, stack_size=self._stack_size)
        return self._internal_env

    def reset(self, deck_state_dict=None):
        self.internal_env.reset(deck_state_dict=deck_state_dict)

    def notify_of_action(self, p_id_acted, action_he_did):
        assert self.internal_env.current_player.seat_id == p_id_acted
        self.internal_env.step(action_he_did)

    def notify_of_raise_frac_action(self, p_id_acted, frac):
        assert self.internal_env.current_player.seat_id == p_id_acted
        self.internal_env.step_raise_pot_frac(pot_frac=frac)

    def env_state_dict(self):
        return self.internal_env.state_dict()

    def load_env_state_dict(self, state_dict):
        self.internal_env.load_state_dict(state_dict)

    def _uniform(self):
        """one uniform random number per sampled action (tests replay recorded ones)"""
        import numpy as np
        return float(np.random.random())

    def get_action(self, step_env=True, need_probs=False):
        """EvalAgentBase.get_action (:52-63): sample the action of the hand the agent holds at its own table from
        get_a_probs_for_each_hand(); optionally step the table; optionally return the whole [RANGE_SIZE, N_ACTIONS] table"""
        import numpy as np
        env = self.internal_env
        probs = np.asarray(self.get_a_probs_for_each_hand())
        row = probs[env.get_range_idx(p_id=env.current_player.seat_id)].astype(np.float64)
        action = int(min(np.searchsorted(np.cumsum(row), self._uniform(), side="right"), row.size - 1))
        while row[action] == 0 and action > 0:  # the draw fell beyond the last action with mass through rounding
            action -= 1
        if step_env:
            env.step(action)
        return action, (probs if need_probs else None)

    def update_weights(self, weights_for_eval_agent):
        raise NotImplementedError

    # ---- state
    def set_stack_size(self, stack_size):
        self._stack_size = stack_size
        self._internal_env = None  # rebuilt with the new stacks on next use

    def get_mode(self):
        return self._mode

    def set_mode(self, mode):
        assert mode in self.ALL_MODES
        self._mode = mode

    def set_to_public_tree_node_state(self, node):
        self._node = node

    def _state_dict(self):
        raise NotImplementedError

    def _load_state_dict(self, state):
        raise NotImplementedError

    def state_dict(self):
        return {"t_prof": self.t_prof, "mode": self._mode, "agent": self._state_dict()}

    def load_state_dict(self, state):
        self._mode = state["mode"]
        self._load_state_dict(state["agent"])

    def store_to_disk(self, path, file_name):
        import os
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, str(file_name) + ".pkl"), "wb") as f:
            pickle.dump(self.state_dict(), f)

    @classmethod
    def load_from_disk(cls, path_to_eval_agent):
        with open(path_to_eval_agent, "rb") as f:
            state = pickle.load(f)
        agent = cls(t_prof=state["t_prof"])
        agent.load_state_dict(state)
        return agent
