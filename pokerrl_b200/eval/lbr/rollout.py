"""LBR roll-outs on the GPU (host side of csrc/lbr_rollout.cu).

`lbr_checkdown_equity` evaluates a BATCH of (LBR hand, dealt board, agent range) queries of one street at once;
`LBRRolloutManager` keeps the call surface of the reference's per-decision helper
(`PokerRL/eval/lbr/LocalLBRWorker.py:377-512`: built from the env's board and the LBR hand, then
`get_lbr_checkdown_equity(agent_range)`), so a worker loop written against the reference can swap it in.
"""
import ctypes as C

import numpy as np
import torch

from pokerrl_b200 import _native as nat
from pokerrl_b200.game.Poker import Poker


def lbr_checkdown_equity(lbr_hands_1d, boards_1d, n_dealt, ranges, device=None, reference_board_counter_quirk=False):
    """lbr_hands_1d int8 [B, 2], boards_1d int8 [B, 5] (dealt cards first), ranges float32 [B, 1326] -> torch float32 [B].
    reference_board_counter_quirk: compare ranks on the first completion for every completion, as the reference's
    _calc_eq does because it never advances `_i` (LocalLBRWorker.py:468-512) - for parity checks only."""
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    h = torch.as_tensor(np.ascontiguousarray(lbr_hands_1d, np.int8)).to(dev)
    b = torch.as_tensor(np.ascontiguousarray(boards_1d, np.int8)).to(dev)
    r = torch.as_tensor(ranges).to(device=dev, dtype=torch.float32).contiguous()
    n = int(h.shape[0])
    assert h.shape == (n, 2) and b.shape == (n, 5) and r.shape == (n, 1326) and 0 <= n_dealt <= 5
    ws = torch.empty(int(nat.lib().prl_lbr_workspace_doubles(n, int(n_dealt))), dtype=torch.float64, device=dev)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.call("prl_lbr_checkdown_equity", C.c_void_p(h.data_ptr()), C.c_void_p(b.data_ptr()), int(n_dealt),
                 C.c_void_p(r.data_ptr()), n, int(reference_board_counter_quirk), C.c_void_p(ws.data_ptr()),
                 C.c_void_p(out.data_ptr()),
                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return out


class LBRRolloutManager:
    """_LBRRolloutManager(t_prof, env_bldr, env, lbr_hand_2d) of the reference: board and street are read from `env`
    (anything with `.board` [5, 2] int8 incl. not-dealt tokens); the per-board rank comparisons the reference precomputes in
    __init__ (:392-424) happen inside the kernel."""

    def __init__(self, t_prof, env_bldr, env, lbr_hand_2d, device=None, reference_board_counter_quirk=False):
        self._quirk = bool(reference_board_counter_quirk)
        lut = env_bldr.lut_holder
        self._hand = np.sort(np.asarray(lut.get_1d_cards(np.asarray(lbr_hand_2d))).reshape(-1)).astype(np.int8)
        board = np.asarray(lut.get_1d_cards(np.asarray(env.board))).reshape(-1).astype(np.int8)
        dealt = board[board != Poker.CARD_NOT_DEALT_TOKEN_1D]
        self._n_dealt = int(dealt.size)
        self._board = np.full(5, Poker.CARD_NOT_DEALT_TOKEN_1D, np.int8)
        self._board[:self._n_dealt] = dealt
        self._device = device

    def get_lbr_checkdown_equity(self, agent_range):
        rng = agent_range.range if hasattr(agent_range, "range") else agent_range
        return float(lbr_checkdown_equity(self._hand[None], self._board[None], self._n_dealt, np.asarray(rng, np.float32)[None],
                                          device=self._device, reference_board_counter_quirk=self._quirk)[0].item())
