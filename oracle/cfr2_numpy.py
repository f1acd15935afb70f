"""CPU restatement (numpy, float64) of the tabular CFR / value path GENERALISED TO TWO-HOLE-CARD GAMES.
TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline).

Parity status: "parity unpinned" w.r.t. the reference - the reference's ValueFiller / PublicTree only work for
one-card games (ValueFiller.py:18-19, PublicTree.py:193-203; SURVEY.md headline 2), so no reference output exists for
Hold'em trees.  What IS pinned: (a) with a one-card rule set this file reproduces oracle/cfr_numpy.py (hence the
reference) to float64 round-off (tests/test_oracle_cfr2.py), (b) hand strengths come from the evaluator pinned to
lib_hand_eval.so, (c) invariants: zero-sum at every node (ValueFiller.py:98), exploitability >= 0 and -> 0,
suit-isomorphic evaluation == full evaluation.

Generalisation of the one-card constants: opponent-hand normaliser K = C(n,2)/C(n-2,2)
(n = deck size) instead of n/(n-1) (ValueFiller.py:19); "other hand" means "shares no card" (fold: inclusion-exclusion
over the two cards; showdown: sign matrix masked by card compatibility); a k-card chance deal multiplies BOTH reach rows
by board_prob (1/C(n-4,k) for the full game) and zeroes hands that hold a board card (StrategyFiller.py:159-166); the
parent of the boards sums board_mult-weighted child values and, with suit isomorphism, symmetrises the sum over the
suit permutations.
"""
import numpy as np

KIND_P0, KIND_P1, KIND_CHANCE, KIND_FOLD, KIND_SHOWDOWN, KIND_SHOWDOWN_ALLIN = 0, 1, 2, 3, 4, 5


def fold_row(inc, ro):
    """ValueFiller.py:103-125 for two-card hands (before K, sign and pot): mass of the opponent hands sharing no card
    with mine = total - (hands holding my first card) - (hands holding my second card) + (my own hand, subtracted twice).
    inc = hand-card incidence [R, n_deck]."""
    cs = inc.T @ ro  # per-card sums
    return ro.sum() - inc @ cs + (int(inc[0].sum()) - 1) * ro


def sign_matrix(ranks, compat, blocked):
    """ValueFiller.py:140-155 as a matrix: S[h, h'] = sign(rank_h - rank_h') for card-disjoint live hands, else 0"""
    rk = np.asarray(ranks).astype(np.int64)
    s = np.sign(rk[:, None] - rk[None, :]).astype(np.float64)
    return s * (compat & ~blocked[:, None] & ~blocked[None, :])


def allin_equity_matrix(ranks, weight, hand_cards, n_deck, sym_perm=None):
    """E[h][h'] = sum over the hand permutations q and the boards b of weight_b * S_b[q(h)][q(h')] with S_b = sign_matrix
    of board b (float64, brute force) = the matrix of ALL boards of the orbits.  The value rows of an all-in showdown before
    the deal are K * pot / 2 * E @ reach_opp - for suit-symmetric reach rows (the isomorphism contract of
    game/holdem_boards.py) exactly what a chance node over those boards with showdown children and board_prob * board_mult =
    weight gives."""
    R = ranks.shape[1]
    inc = np.zeros((R, n_deck))
    for k in range(hand_cards.shape[1]):
        inc[np.arange(R), hand_cards[:, k]] = 1
    compat = (inc @ inc.T) == 0
    ec = np.zeros((R, R))
    for b in range(ranks.shape[0]):
        ec += weight[b] * sign_matrix(ranks[b], compat, ranks[b] < 0)
    if sym_perm is None:
        return ec
    return sum(ec[np.ix_(np.asarray(pm, np.int64), np.asarray(pm, np.int64))] for pm in sym_perm)


class Oracle2Tree:
    def __init__(self, ft, hand_cards, board_ranks, board_prob, board_mult, sym_perm=None, eq_const=None):
        """ft: FlatTree; hand_cards int[R, n_hole]; board_ranks int32[n_boards_total, R] (global board id order,
        -1 = blocked / board incomplete); board_prob / board_mult float[n_boards_total]; sym_perm int[n_sym, R]."""
        self.ft = ft
        self.R = R = ft.R
        self.N = ft.n_nodes
        self.hand_cards = np.asarray(hand_cards)
        n_deck = ft.rules.N_CARDS_IN_DECK
        n_hole = self.hand_cards.shape[1]
        if eq_const is None:
            from math import comb
            eq_const = comb(n_deck, n_hole) / comb(n_deck - n_hole, n_hole)
        self.K = eq_const
        self.board_ranks = np.asarray(board_ranks)
        self.board_prob = np.asarray(board_prob, np.float64)
        self.board_mult = np.asarray(board_mult, np.float64)
        self.sym_perm = None if sym_perm is None else np.asarray(sym_perm)
        # hand incidence [R, n_deck] and compatibility [R, R]
        inc = np.zeros((R, n_deck))
        for k in range(n_hole):
            inc[np.arange(R), self.hand_cards[:, k]] = 1
        self.inc = inc
        self.compat = (inc @ inc.T) == 0
        bc = ft.board_cards()  # [n_boards_total, n_board_cards], -127 padded
        self.board_blocked = np.zeros((bc.shape[0], R), bool)
        for b in range(bc.shape[0]):
            cards = bc[b][bc[b] >= 0]
            if cards.size:
                self.board_blocked[b] = inc[:, cards].sum(axis=1) > 0
        self.reach = np.zeros((self.N, 2, R))
        self.ev = np.zeros((self.N, 2, R))
        self.ev_br = np.zeros((self.N, 2, R))
        self.strategy = [None] * self.N
        self._sign = {}
        self.allin_equity = None  # float64 [R, R] (or {board id: matrix}), see allin_equity_matrix()

    def decision_nodes(self):
        ft = self.ft
        return np.nonzero((ft.kind <= KIND_P1) & (ft.first_child >= 0))[0]

    def fill_uniform(self):
        for n in self.decision_nodes():
            a = int(self.ft.n_children[n])
            self.strategy[n] = np.full((self.R, a), 1.0 / a)
        self.update_reach()

    def update_reach(self):
        ft = self.ft
        self.reach[0] = 1.0 / self.R
        if ft.board[0] >= 0:  # sub-game root that already shows a board: blocked hands cannot be held
            self.reach[0][:, self.board_blocked[ft.board[0]]] = 0.0
        for n in range(self.N):
            nc = ft.n_children[n]
            if nc == 0:
                continue
            fc, k = ft.first_child[n], ft.kind[n]
            if k <= KIND_P1:
                for a in range(nc):
                    self.reach[fc + a] = self.reach[n]
                    self.reach[fc + a, k] = self.strategy[n][:, a] * self.reach[n, k]
            else:
                for c in range(nc):
                    b = ft.board[fc + c]
                    self.reach[fc + c] = self.reach[n] * np.where(self.board_blocked[b], 0.0, self.board_prob[b])

    def _sign_matrix(self, b):
        if b not in self._sign:
            self._sign[b] = sign_matrix(self.board_ranks[b], self.compat, self.board_blocked[b])
            if len(self._sign) > 64:
                self._sign.pop(next(iter(self._sign)))
        return self._sign[b]

    def compute_ev(self):
        ft = self.ft
        for n in range(self.N - 1, -1, -1):
            k, nc = ft.kind[n], ft.n_children[n]
            if k >= KIND_FOLD:
                b = ft.board[n]
                eq = np.zeros((2, self.R))
                for p in range(2):
                    ro = self.reach[n, 1 - p]
                    if k == KIND_FOLD:
                        e = fold_row(self.inc, ro)  # opponent hands sharing no card with mine
                        eq[p] = -e if ft.acted_last[n] == p else e
                    elif k == KIND_SHOWDOWN:
                        eq[p] = self._sign_matrix(b) @ ro
                    else:  # all-in before the deal: the showdown rows of every board it runs out over, summed like a
                        # chance node's children would be (ValueFiller.py:160-175 is the one-card analogue)
                        if self.allin_equity is None:
                            raise NotImplementedError("all-in showdown before the board is complete: set allin_equity")
                        E = self.allin_equity[int(b)] if isinstance(self.allin_equity, dict) else self.allin_equity
                        eq[p] = E @ ro
                eq *= self.K
                if b >= 0:
                    eq[:, self.board_blocked[b]] = 0.0
                self.ev[n] = eq * float(ft.pot[n]) / 2
                self.ev_br[n] = self.ev[n]
                continue
            fc = ft.first_child[n]
            ev_all, evbr_all = self.ev[fc:fc + nc], self.ev_br[fc:fc + nc]
            if k == KIND_CHANCE:
                w = self.board_mult[ft.board[fc:fc + nc]][:, None, None]
                e, eb = (w * ev_all).sum(axis=0), (w * evbr_all).sum(axis=0)
                if self.sym_perm is not None:
                    e = sum(e[:, pm] for pm in self.sym_perm)
                    eb = sum(eb[:, pm] for pm in self.sym_perm)
                self.ev[n], self.ev_br[n] = e, eb
            else:
                p, o = int(k), 1 - int(k)
                self.ev[n, p] = (self.strategy[n].T * ev_all[:, p]).sum(axis=0)
                self.ev[n, o] = ev_all[:, o].sum(axis=0)
                self.ev_br[n, o] = evbr_all[:, o].sum(axis=0)
                self.ev_br[n, p] = evbr_all[:, p].max(axis=0)
        self.exploitability = ((self.ev_br[0] - self.ev[0]) * self.reach[0]).sum(axis=1)
        return self.exploitability


class Oracle2CFR:
    """_CFRBase / CFRPlus / LinearCFR / VanillaCFR (PokerRL/cfr/) in float64 on an Oracle2Tree."""

    def __init__(self, tree, algo="CFRPlus", delay=0, ev_normalizer=1.0):
        self.t, self.algo, self.delay = tree, algo, delay
        self.ft, self.R = tree.ft, tree.R
        self.ev_normalizer = ev_normalizer
        self.iter_counter = 0
        N = self.ft.n_nodes
        self.regret, self.avg, self.avg_sum = [None] * N, [None] * N, [None] * N
        self.t.fill_uniform()

    def _nodes_of(self, p):
        ft = self.ft
        return np.nonzero((ft.kind == p) & (ft.first_child >= 0))[0]

    def iteration(self):
        t, ft = self.t, self.ft
        for p in range(2):
            t.compute_ev()
            for n in self._nodes_of(p):
                fc, A = ft.first_child[n], int(ft.n_children[n])
                d = t.ev[fc:fc + A, p].T - t.ev[n, p][:, None]
                old = self.regret[n] if self.regret[n] is not None else np.zeros((self.R, A))
                if self.algo == "CFRPlus":
                    reg = np.maximum(d + old, 0)
                elif self.algo == "LinearCFR":
                    reg = (self.iter_counter + 1) * d + old
                else:
                    reg = d + old
                self.regret[n] = reg
                rp = np.maximum(reg, 0)
                s = rp.sum(axis=1, keepdims=True)
                with np.errstate(divide="ignore", invalid="ignore"):
                    t.strategy[n] = np.where(s > 0, rp / s, 1.0 / A)
            t.update_reach()
            for n in self._nodes_of(p):
                A = int(ft.n_children[n])
                if self.algo == "CFRPlus":
                    if self.iter_counter > self.delay:
                        cw = sum(range(self.delay + 1, self.iter_counter + 1))
                        nw = self.iter_counter - self.delay + 1
                        self.avg[n] = cw / (cw + nw) * self.avg[n] + nw / (cw + nw) * t.strategy[n]
                    elif self.iter_counter == self.delay:
                        self.avg[n] = t.strategy[n].copy()
                else:
                    c = t.strategy[n] * t.reach[n, p][:, None]
                    if self.algo == "LinearCFR":
                        c = c * (self.iter_counter + 1)
                    self.avg_sum[n] = c if self.avg_sum[n] is None else self.avg_sum[n] + c
                    s = self.avg_sum[n].sum(axis=1, keepdims=True)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        self.avg[n] = np.where(s == 0, 1.0 / A, self.avg_sum[n] / s)
        self.iter_counter += 1

    def _metric(self, expl):
        return float(sum(expl[p] * self.ev_normalizer for p in range(2)) / 2)

    def exploitability_current(self):
        return self._metric(self.t.compute_ev())

    def exploitability_average(self):
        keep = self.t.strategy
        self.t.strategy = [None if a is None else a.copy() for a in self.avg]
        self.t.update_reach()
        e = self._metric(self.t.compute_ev())
        self.t.strategy = keep
        self.t.update_reach()
        return e
