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

    def update_weights(self, weights_for_eval_agent):
        raise NotImplementedError

    # ---- state
    def set_stack_size(self, stack_size):
        self._stack_size = stack_size

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
