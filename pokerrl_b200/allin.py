"""Equity matrix of an all-in showdown before the board is complete (two-card games) and its dense product with reach
rows on the tensor cores - host side of csrc/allin_dense.cu.

Reference: ValueFiller.py:160-175 (`_get_call_eq_preflop`) enumerates the missing board per terminal for one-card games;
for two-card hands the sum over the boards is strategy-independent and is folded into one matrix
    E[h][h'] = sum_q sum_b prob_b * mult_b * sign(rank_b(q(h)) - rank_b(q(h')))
(q: suit permutations of the isomorphism contract, game/holdem_boards.py).  No CPU fallback: every step is a kernel."""
import ctypes as C

import numpy as np
import torch

from pokerrl_b200 import _native as nat


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class AllinEquity:
    def __init__(self, rules, spec=None, device=None, ranks=None, chunk=16384, boards=None, weights=None, sym_perm=None,
                 keep_ec=False):
        """spec: holdem_boards.BoardSpec (boards, board_prob, board_mult, sym_perm) - or boards int8 [n, 5], weights
        float64 [n] (deal probability x weight in the parent's sum) and sym_perm given directly; ranks: optional DEVICE
        int32 [n_boards, R] hand strengths of the boards (else computed here by prl_hand_rank_boards)."""
        from pokerrl_b200.hand_eval import hand_rank_all_hands_on_given_boards
        from pokerrl_b200.solver import _require_cuda
        self.device = dev = _require_cuda(device)
        self.R = R = rules.RANGE_SIZE
        lut = rules.get_lut_holder()
        if spec is not None:
            boards, sym_perm = spec.boards, spec.sym_perm
            weights = np.asarray(spec.board_prob, np.float64) * np.asarray(spec.board_mult, np.float64)
        boards = np.ascontiguousarray(boards, np.int8)
        w = np.ascontiguousarray(weights, np.float64)
        with torch.cuda.device(dev):
            self.hand_cards = torch.from_numpy(np.ascontiguousarray(lut.LUT_IDX_2_HOLE_CARDS, np.int8)).to(dev)
            ec = torch.zeros(R, R, dtype=torch.float64, device=dev)
            wt = torch.from_numpy(w).to(dev)
            for i in range(0, boards.shape[0], chunk):
                rk = ranks[i:i + chunk] if ranks is not None else hand_rank_all_hands_on_given_boards(boards[i:i + chunk], device=dev)
                rk = rk.contiguous()
                n = int(rk.shape[0])
                nat.call("prl_allin_equity_accumulate", C.c_void_p(rk.data_ptr()), C.c_void_p(wt[i:i + n].data_ptr()), n, R,
                         C.c_void_p(ec.data_ptr()), _stream(dev))
            sp = sym_perm
            self.sym = torch.from_numpy(np.ascontiguousarray(sp, np.int16)).to(dev) if sp is not None else None
            self.tiles = torch.zeros(int(nat.lib().prl_allin_tiles_bytes(R)), dtype=torch.uint8, device=dev)
            self.partial = torch.zeros(int(nat.lib().prl_allin_partial_bytes(R)) // 4, dtype=torch.float32, device=dev)
            nat.call("prl_allin_equity_finish", C.c_void_p(ec.data_ptr()), R, C.c_void_p(self.hand_cards.data_ptr()),
                     C.c_void_p(self.sym.data_ptr()) if self.sym is not None else None,
                     int(sp.shape[0]) if sp is not None else 0, C.c_void_p(self.tiles.data_ptr()), _stream(dev))
            self.ec = ec if keep_ec else None  # unsymmetrised float64 sums (14 MB): kept on request (inspection / tests)

    def values(self, x, scale=None, out=None):
        """y[c] = scale[c] * E @ x[c] for the rows of x (float32 [n_cols, ld >= R], device) - one tensor-core launch per 16 rows"""
        assert x.dtype == torch.float32 and x.is_cuda and x.dim() == 2 and x.shape[1] >= self.R and x.is_contiguous()
        n = int(x.shape[0])
        y = out if out is not None else torch.zeros_like(x)
        sc = np.ones(n, np.float32) if scale is None else np.ascontiguousarray(scale, np.float32)
        ptr = C.c_void_p * n
        xs = ptr(*[x[c].data_ptr() for c in range(n)])
        ys = ptr(*[y[c].data_ptr() for c in range(n)])
        with torch.cuda.device(self.device):
            nat.call("prl_allin_values", C.c_void_p(self.tiles.data_ptr()), self.R, xs, ys, None, sc.ctypes.data_as(C.c_void_p), n,
                     C.c_void_p(self.partial.data_ptr()), _stream(self.device))
        return y
