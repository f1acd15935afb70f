"""Host restatement of the O(R) showdown / fold evaluation that terminal2_kernel performs (pokerrl_b200/csrc/
cfr_twocard.cu), checked against the O(R^2) definition (oracle/cfr2_numpy.py uses the same
definition).  It follows the kernel's data flow step by step - strength-order tables as prl_board_order_tables builds them,
quad-wise card-row scans, centred prefix sums E[i] = prefix(i) - total / 2, the packed per-hand record - so an indexing
mistake in that scheme shows up here, without a GPU."""
import numpy as np
import pytest

from twocard_common import oracle_ranks

N_DECK, ROW_STRIDE = 52, 53
C1, C2 = np.triu_indices(N_DECK, k=1)
R = C1.size


def pair_index(a, b):
    c1, c2 = min(a, b), max(a, b)
    return c1 * (2 * N_DECK - 1 - c1) // 2 + (c2 - c1 - 1)


def order_tables(ranks):
    """board_order_kernel + board_rows_kernel (ranks int32 [R], -1 = blocked)"""
    live = ranks >= 0
    gs, ge, pos = (np.full(R, -1, np.int64) for _ in range(3))
    lr = ranks[live]
    for h in np.nonzero(live)[0]:
        gs[h] = (lr < ranks[h]).sum()
        ge[h] = (lr <= ranks[h]).sum()
        pos[h] = gs[h] + ((ranks[:h] == ranks[h]) & live[:h]).sum()
    row_len = N_DECK - 1
    row_order = np.full((N_DECK, row_len), -1, np.int64)
    row_pos = np.zeros((R, 4), np.int64)
    for cc in range(N_DECK):
        hands = [pair_index(cc, j + (j >= cc)) for j in range(row_len)]
        g = np.array([gs[h] for h in hands])
        for j, h in enumerate(hands):
            if g[j] < 0:
                continue
            ok = g >= 0
            lt = (ok & (g < g[j])).sum()
            le = (ok & (g <= g[j])).sum()
            tb = (ok[:j] & (g[:j] == g[j])).sum()
            row_order[cc, lt + tb] = h
            k = 0 if cc == C1[h] else 1
            row_pos[h, k], row_pos[h, 2 + k] = lt, le
    return gs, ge, pos, row_order, row_pos


def hand_rec(gs, ge, row_pos):
    rec = np.zeros((R, 8), np.int64)
    rec[:, 0], rec[:, 1] = gs, ge
    rec[:, 2], rec[:, 3] = C1 * ROW_STRIDE + row_pos[:, 0], C1 * ROW_STRIDE + row_pos[:, 2]
    rec[:, 4], rec[:, 5] = C2 * ROW_STRIDE + row_pos[:, 1], C2 * ROW_STRIDE + row_pos[:, 3]
    return rec


def showdown_kernel_flow(ro, tabs, n_threads=256):
    gs, ge, pos, row_order, row_pos = tabs
    row_len, seg = N_DECK - 1, (N_DECK - 1 + 3) >> 2
    srt = np.zeros(R + 1)
    for h in range(R):
        if pos[h] >= 0:
            srt[pos[h]] = ro[h]
    rp = np.full(N_DECK * ROW_STRIDE, np.nan)
    for cc in range(N_DECK):  # one quad per card row
        run, inc = np.zeros(4), np.zeros((4, 16))
        for qj in range(4):
            for i in range(16):
                idx = qj * seg + i
                v = 0.0
                if i < seg and idx < row_len and row_order[cc, idx] >= 0:
                    v = ro[row_order[cc, idx]]
                run[qj] += v
                inc[qj, i] = run[qj]
        sc = np.cumsum(run)
        half = 0.5 * sc[3]
        rp[cc * ROW_STRIDE] = -half
        for qj in range(4):
            off = (sc[qj] - run[qj]) - half
            for i in range(16):
                idx = qj * seg + i
                if i < seg and idx < row_len:
                    rp[cc * ROW_STRIDE + idx + 1] = off + inc[qj, i]
    per = (R + 1 + n_threads - 1) // n_threads
    loc = np.array([srt[t * per:min(R + 1, t * per + per)].sum() for t in range(n_threads)])
    excl = np.cumsum(loc) - loc
    total = loc.sum()
    out = srt.copy()
    for t in range(n_threads):
        run = excl[t] - 0.5 * total
        for i in range(t * per, min(R + 1, t * per + per)):
            x = srt[i]
            out[i] = run
            run += x
    srt = out
    rec = hand_rec(gs, ge, row_pos)
    v_rec, v_tab = np.zeros(R), np.zeros(R)
    for h in range(R):
        if gs[h] < 0:
            continue
        q = rec[h]
        v_rec[h] = (srt[q[0]] + srt[q[1]]) - ((rp[q[2]] + rp[q[3]]) + (rp[q[4]] + rp[q[5]]))
        r1, r2 = C1[h] * ROW_STRIDE, C2[h] * ROW_STRIDE
        v_tab[h] = (srt[gs[h]] + srt[ge[h]]) - ((rp[r1 + row_pos[h, 0]] + rp[r1 + row_pos[h, 2]])
                                                + (rp[r2 + row_pos[h, 1]] + rp[r2 + row_pos[h, 3]]))
    return v_rec, v_tab


def fold_kernel_flow(ro):
    row_len, seg = N_DECK - 1, (N_DECK - 1 + 3) >> 2
    cs = np.zeros(N_DECK)
    for cc in range(N_DECK):
        for qj in range(4):
            for i in range(seg):
                idx = qj * seg + i
                if idx < row_len:
                    cs[cc] += ro[pair_index(cc, idx + (idx >= cc))]
    return ro.sum() - cs[C1] - cs[C2] + ro


def fold2_kernel_flow(ro):
    """fold2_kernel: thread (card cc, quarter j) adds every fourth hand (r < cc, cc) and every fourth hand (cc, x > cc),
    addressed through the lexicographic layout of the range; four partials per card"""
    base = [c * (2 * N_DECK - 1 - c) // 2 for c in range(64)]
    part = np.zeros((4, 64))
    for j in range(4):
        for cc in range(N_DECK):
            acc = 0.0
            for r in range(j, cc, 4):
                acc += ro[base[r] + cc - r - 1]
            for k in range(j, N_DECK - 1 - cc, 4):
                acc += ro[base[cc] + k]
            part[j, cc] = acc
    cs = (part[0] + part[1]) + (part[2] + part[3])
    return ro.sum() - cs[C1] - cs[C2] + ro


@pytest.mark.parametrize("seed", [0, 1])
def test_showdown_flow_equals_definition(seed):
    rng = np.random.default_rng(seed)
    board = np.sort(rng.choice(N_DECK, 5, replace=False))
    if seed == 1:
        board = np.array([0, 4, 8, 12, 17])  # straight on the board: large tie groups
    ranks = oracle_ranks(board[None, :])[0].astype(np.int64)
    blocked = np.isin(C1, board) | np.isin(C2, board)
    assert ((ranks < 0) == blocked).all()
    ro = rng.random(R) * ~blocked
    tabs = order_tables(ranks)
    v_rec, v_tab = showdown_kernel_flow(ro, tabs)
    share = (C1[:, None] == C1[None, :]) | (C1[:, None] == C2[None, :]) | (C2[:, None] == C1[None, :]) | \
            (C2[:, None] == C2[None, :])
    sign = np.sign(ranks[:, None] - ranks[None, :]) * ~share * ~blocked[None, :] * ~blocked[:, None]
    want = sign @ ro
    np.testing.assert_allclose(v_rec, want, atol=1e-9)
    np.testing.assert_allclose(v_tab, want, atol=1e-9)


def test_fold_flow_equals_definition():
    rng = np.random.default_rng(5)
    ro = rng.random(R)
    share = (C1[:, None] == C1[None, :]) | (C1[:, None] == C2[None, :]) | (C2[:, None] == C1[None, :]) | \
            (C2[:, None] == C2[None, :])
    np.testing.assert_allclose(fold_kernel_flow(ro), (~share).astype(float) @ ro, atol=1e-9)
    np.testing.assert_allclose(fold2_kernel_flow(ro), (~share).astype(float) @ ro, atol=1e-9)
