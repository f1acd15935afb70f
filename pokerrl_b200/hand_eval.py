"""Batched Hold'em hand evaluation on the GPU (host side of csrc/hand_eval.cu).

Mirrors `HoldemRules.get_hand_rank_all_hands_on_given_boards` / `get_hand_rank` (`PokerRL/game/_/rl_env/game_rules.py:
213-223`) with torch tensors as buffers; values are bit-identical to the reference's lib_hand_eval.so."""
import ctypes as C

import numpy as np
import torch

from pokerrl_b200 import _native as nat


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def hand_rank_all_hands_on_given_boards(boards_1d, device=None):
    """boards_1d: int8 [N, 5] (numpy or torch) -> torch int32 [N, 1326] on the device (-1 = blocked by the board)."""
    b = torch.as_tensor(np.ascontiguousarray(boards_1d) if isinstance(boards_1d, np.ndarray) else boards_1d)
    device = device if device is not None else "cuda:%d" % torch.cuda.current_device()
    b = b.to(device=device, dtype=torch.int8).contiguous()
    assert b.dim() == 2 and b.shape[1] == 5
    out = torch.empty((b.shape[0], 1326), dtype=torch.int32, device=b.device)
    with torch.cuda.device(b.device):
        nat.call("prl_hand_rank_boards", C.c_void_p(b.data_ptr()), int(b.shape[0]), C.c_void_p(out.data_ptr()), _stream(b.device))
    return out


def hand_rank_7(cards_1d, device=None):
    """cards_1d: int8 [N, 7] -> torch int32 [N]"""
    c = torch.as_tensor(np.ascontiguousarray(cards_1d) if isinstance(cards_1d, np.ndarray) else cards_1d)
    device = device if device is not None else "cuda:%d" % torch.cuda.current_device()
    c = c.to(device=device, dtype=torch.int8).contiguous()
    assert c.dim() == 2 and c.shape[1] == 7
    out = torch.empty((c.shape[0],), dtype=torch.int32, device=c.device)
    with torch.cuda.device(c.device):
        nat.call("prl_hand_rank_7", C.c_void_p(c.data_ptr()), int(c.shape[0]), C.c_void_p(out.data_ptr()), _stream(c.device))
    return out
