"""Local Best Response (Lisy & Bowling, arXiv:1612.07547) against an eval agent - the episode loop of
`PokerRL/eval/lbr/LocalLBRWorker.py:12-308` on the device engine.

LBR plays one seat of a heads-up hand.  At each of its decisions it tracks the agent's range (PokerRange), and for every
action it may take it estimates the value of taking it and then check / calling to showdown:
    fold 0;  call  wp * pot - (1 - wp) * asked;  raise r  p_fold * pot + (1 - p_fold) * (wp_r * pot_r - (1 - wp_r) * put_in_r)
with wp = the roll-out equity of its hand against the agent's range over the boards still to come, p_fold and the agent's
range after not folding taken from the agent's policy in the state after the raise (LocalLBRWorker.py:93-146, 205-266).
It takes the best one, the hand is played out, LBR's winnings are a lower bound on the agent's exploitability.

What differs from the reference's two near-identical loops (`_run_limit`, `_run_no_limit`):
  * one loop for both betting structures (the fixed-limit raise is "the one raise"),
  * the roll-outs of ALL candidate actions of a decision are ONE batched launch of csrc/lbr_rollout.cu (the reference builds
    a manager that enumerates and ranks every board completion per decision on the host),
  * both tables (LBR's and the agent's) are single-table views of the device engine (game/poker_env.py).
`reference_board_counter_quirk=True` reproduces the reference's roll-out defect (ranks of the first completion for every
completion, rollout.py) so that its episodes can be replayed exactly (tests/test_gpu_lbr_worker.py)."""
import numpy as np

from pokerrl_b200.eval.lbr.rollout import lbr_checkdown_equity
from pokerrl_b200.game.Poker import Poker
from pokerrl_b200.game.PokerRange import PokerRange


class LocalLBRWorker:
    def __init__(self, t_prof, chief_handle, eval_agent_cls, reference_board_counter_quirk=False, device=None):
        assert t_prof.n_seats == 2
        self.t_prof, self.chief_handle = t_prof, chief_handle
        self.lbr_args = t_prof.module_args["lbr"]
        self.check_to_round = self.lbr_args.lbr_check_to_round
        self.agent = eval_agent_cls(t_prof=t_prof)
        self._env_bldr = self.agent.env_bldr  # LBR sits at a table with the agent's bet sizes (see step_raise_pot_frac)
        lbr_set = getattr(self.lbr_args, "lbr_bet_set", None)
        mine = getattr(self._env_bldr.env_args, "bet_sizes_list_as_frac_of_pot", None)
        if self._env_bldr.env_cls.BETTING == "discretized" and lbr_set is not None and sorted(lbr_set) != sorted(mine):
            raise NotImplementedError("LBR with a bet set other than the agent's needs raises by arbitrary pot fractions at the "
                                      "agent's table; the device engine steps discrete actions only")
        assert self.check_to_round is None or self.check_to_round in self._env_bldr.rules.ALL_ROUNDS_LIST
        self._quirk, self._device = bool(reference_board_counter_quirk), device
        self.agent_range = PokerRange(env_bldr=self._env_bldr)
        self._env = None
        self.last_utilities = []  # utility vectors of the hand played last (diagnostics / parity tests)

    # ------------------------------------------------------------------------------------------------------------ API
    def run(self, agent_seat_id, n_iterations, mode, stack_size, decks=None):
        """float32 [n_iterations]: LBR's winnings per hand in the game's EV unit (LocalLBRWorker.py:36-50); None if the agent
        cannot be evaluated in `mode` yet.  decks: optional list of deck_state_dicts to deal from (replays)."""
        self.agent.set_mode(mode)
        self.agent.set_stack_size(stack_size)
        self.agent_range.reset()
        self._env = self._env_bldr.get_new_env(is_evaluating=True, stack_size=stack_size)
        if not self.agent.can_compute_mode():
            return None
        out = np.empty(n_iterations, dtype=np.float32)
        for i in range(n_iterations):
            out[i] = self.play_hand(agent_seat_id, None if decks is None else decks[i])
        return out

    def update_weights(self, weights_for_eval_agent):
        self.agent.update_weights(weights_for_eval_agent)

    # ------------------------------------------------------------------------------------------------------ one hand
    def play_hand(self, agent_seat_id, deck_state_dict=None):
        env, agent, rng = self._env, self.agent, self.agent_range
        lbr_seat = 1 - agent_seat_id
        _, reward, done, _ = env.reset(deck_state_dict=deck_state_dict)
        agent.reset(deck_state_dict=env.cards_state_dict() if deck_state_dict is None else deck_state_dict)
        rng.reset()
        lbr_hand = env.get_hole_cards_of_player(p_id=lbr_seat)
        rng.set_cards_to_zero_prob(cards_2d=lbr_hand)
        self.last_utilities = []
        while not done:
            raise_frac = None
            if env.current_player.seat_id == lbr_seat:
                if self.check_to_round is not None and env.current_round < self.check_to_round:
                    action = Poker.CHECK_CALL
                else:
                    action = self._best_action(agent_seat_id, lbr_hand)
                if action >= 2 and env.bet_sizes_list_as_frac_of_pot is not None:
                    raise_frac = env.bet_sizes_list_as_frac_of_pot[action - 2]
                    agent.notify_of_raise_frac_action(p_id_acted=lbr_seat, frac=raise_frac)
                else:
                    agent.notify_of_action(p_id_acted=lbr_seat, action_he_did=action)
            else:
                action, probs = agent.get_action(step_env=True, need_probs=True)
                rng.update_after_action(action=action, all_a_probs_for_all_hands=probs)
                if action >= 2 and env.bet_sizes_list_as_frac_of_pot is not None:  # the size is the AGENT's table's
                    raise_frac = sorted(agent.env_bldr.env_args.bet_sizes_list_as_frac_of_pot)[action - 2]
            round_before = env.current_round
            if raise_frac is not None:
                _, reward, done, _ = env.step_raise_pot_frac(pot_frac=raise_frac)
            else:
                _, reward, done, _ = env.step(action)
            if env.current_round != round_before:
                rng.update_after_new_round(new_round=env.current_round, board_now_2d=env.board)
        return reward[lbr_seat] * env.REWARD_SCALAR * env.EV_NORMALIZER

    # ----------------------------------------------------------------------------------------------- LBR's decision
    def _best_action(self, agent_seat_id, lbr_hand_2d):
        env, agent, rng = self._env, self.agent, self.agent_range
        lbr_seat = 1 - agent_seat_id
        n_actions = 3 if env.bet_sizes_list_as_frac_of_pot is None else 2 + len(env.bet_sizes_list_as_frac_of_pot)
        utility = np.full(n_actions, -1.0, dtype=np.float32)  # illegal: -1, fold: 0 (LocalLBRWorker.py:94-98)
        utility[Poker.FOLD] = 0.0
        seats = env.seats
        asked = seats[agent_seat_id].current_bet - seats[lbr_seat].current_bet
        pot_now = env.get_all_winnable_money()
        raises = [a for a in env.get_legal_actions() if a >= 2]
        # every candidate's agent range first (host logic on the two tables), then ONE launch for all their roll-outs
        ranges, after_raise = [np.copy(rng.range)], []
        if raises:
            saved_env, saved_agent_env, saved_range = env.state_dict(), agent.env_state_dict(), rng.state_dict()
            for a in raises:
                env.step(a)
                pot_after = env.get_all_winnable_money()
                if env.bet_sizes_list_as_frac_of_pot is not None:
                    agent.notify_of_raise_frac_action(p_id_acted=lbr_seat, frac=env.bet_sizes_list_as_frac_of_pot[a - 2])
                else:
                    agent.notify_of_action(p_id_acted=lbr_seat, action_he_did=a)
                # the agent's answer to the raise; its table is not stepped.  (In fixed-limit games the reference asks through
                # get_action(step_env=False), which samples and discards an action - one random number; kept, so that a
                # recorded episode of the reference replays draw for draw.)
                if env.IS_FIXED_LIMIT_GAME:
                    probs = np.asarray(agent.get_action(step_env=False, need_probs=True)[1])
                else:
                    probs = np.asarray(agent.get_a_probs_for_each_hand())
                fold_prob = np.sum(rng.range * probs[:, Poker.FOLD])
                rng.mul_and_norm(1 - probs[:, Poker.FOLD])
                ranges.append(np.copy(rng.range))
                after_raise.append((a, pot_after, fold_prob))
                rng.load_state_dict(saved_range)
                env.load_state_dict(saved_env)
                agent.load_env_state_dict(saved_agent_env)
        wp = self._rollouts(lbr_hand_2d, ranges)
        utility[Poker.CHECK_CALL] = wp[0] * pot_now - (1 - wp[0]) * asked
        for k, (a, pot_after, fold_prob) in enumerate(after_raise):
            put_in = pot_after - pot_now
            ev_called = wp[1 + k] * pot_after - (1 - wp[1 + k]) * put_in
            utility[a] = fold_prob * pot_now + (1 - fold_prob) * ev_called
        self.last_utilities.append(np.copy(utility))
        return int(np.argmax(utility))

    def _rollouts(self, lbr_hand_2d, ranges):
        """check-down equities of LBR's hand against each of the ranges on the current board: one batched launch"""
        lut = self._env_bldr.lut_holder
        hand = np.sort(np.asarray(lut.get_1d_cards(np.asarray(lbr_hand_2d))).reshape(-1)).astype(np.int8)
        board = np.asarray(lut.get_1d_cards(np.asarray(self._env.board))).reshape(-1).astype(np.int8)
        dealt = board[board != Poker.CARD_NOT_DEALT_TOKEN_1D]
        b = np.full(5, Poker.CARD_NOT_DEALT_TOKEN_1D, np.int8)
        b[:dealt.size] = dealt
        n = len(ranges)
        eq = lbr_checkdown_equity(np.repeat(hand[None], n, axis=0), np.repeat(b[None], n, axis=0), int(dealt.size),
                                  np.asarray(ranges, np.float32), device=self._device, reference_board_counter_quirk=self._quirk)
        return eq.cpu().numpy().astype(np.float32)
