"""Public-tree compiler: betting machine -> abstract betting tree -> depth-sorted SoA arrays ("flat tree").

The reference builds its tree by deep-copying env state dicts while stepping a PokerEnv through every legal
action and every board (`PokerRL/game/_/tree/PublicTree.py:111-293`).  Betting legality never depends on cards,
so here the betting machine (`hu_engine.HUBetting`) is enumerated ONCE into an *abstract* tree whose chance nodes
have a single "any board" child, and the abstract tree is then expanded over the board tables with vectorised
index arithmetic.  The result has exactly the reference's nodes, child order (`allowed_actions` ascending,
PublicTree.py:239; boards in ascending 1D-card order, :193-203) and pots, laid out for the GPU:

  * nodes sorted by depth (level-synchronous sweeps), `level_start[d]` .. `level_start[d+1]`
  * the children of a node are CONTIGUOUS (`first_child`, `n_children`) so child vectors coalesce
  * every child of a decision node owns one row ("slot") of the regret / strategy / average-strategy tables;
    the rows of one decision node are contiguous (`first_slot`)
  * `dfs[n]` = index of node n in the reference's DFS pre-order (used only by parity tests / exports)

Node kinds (`kind`): 0/1 = player 0/1 acts, 2 = chance acts next (reference: PlayerActionNode with
p_id_acting_next == "Ch"), 3 = fold terminal, 4 = showdown on the final street, 5 = all-in showdown before the
board is complete (`ValueFiller.py:34-62`).
"""
from itertools import combinations

import numpy as np

from pokerrl_b200.game import hu_engine as eng
from pokerrl_b200.game.Poker import Poker

KIND_P0, KIND_P1, KIND_CHANCE, KIND_FOLD, KIND_SHOWDOWN, KIND_SHOWDOWN_ALLIN = 0, 1, 2, 3, 4, 5


class _Abs:
    __slots__ = ("kind", "acted_last", "action", "pot", "round", "stack", "bet", "children", "parent", "depth",
                 "cdepth", "allowed")


def enumerate_betting_tree(betting, root_state=None, stop_at_street=None):
    """DFS over the card-free betting machine. Returns the list of abstract nodes (index 0 = root)."""
    nodes = []
    stop = (betting.last_round + 1) if stop_at_street is None else stop_at_street

    def new(kind, acted_last, action, st, parent, depth, cdepth):
        n = _Abs()
        n.kind, n.acted_last, n.action = kind, acted_last, action
        n.pot, n.round = st.main_pot, st.round
        n.stack, n.bet = tuple(st.stack), tuple(st.bet)
        n.children, n.parent, n.depth, n.cdepth, n.allowed = [], parent, depth, cdepth, []
        nodes.append(n)
        return len(nodes) - 1

    s0 = betting.reset() if root_state is None else root_state.copy()
    root = new(s0.cur, -2, -1, s0, -1, 0, 0)
    work = [(root, s0)]
    while work:
        i, st = work.pop()
        me = nodes[i]
        if stop <= st.round:
            continue
        me.allowed = betting.legal_actions(st)
        pend = []
        for a in me.allowed:
            s = st.copy()
            out, pre = betting.step(s, a)
            if out == eng.CONTINUE:
                c = new(s.cur, st.cur, a, s, i, me.depth + 1, me.cdepth)
                pend.append((c, s))
            elif out == eng.NEXT_ROUND:
                # "Ch" node: state before money moves, round still the parent's (PublicTree.py:254-267)
                c = new(KIND_CHANCE, st.cur, a, pre, i, me.depth + 1, me.cdepth)
                # its single abstract child = first node of the new round (one per board after expansion)
                cc = new(s.cur, -1, -1, s, c, me.depth + 2, me.cdepth + 1)
                nodes[c].children.append(cc)
                pend.append((cc, s))
            else:
                kind = {eng.FOLDED: KIND_FOLD, eng.SHOWDOWN: KIND_SHOWDOWN, eng.ALLIN_RUNDOWN: KIND_SHOWDOWN_ALLIN}[out]
                if kind == KIND_SHOWDOWN_ALLIN and st.round == betting.last_round:
                    kind = KIND_SHOWDOWN  # both all-in on the final street: plain showdown
                s.round = st.round  # terminal keeps the parent's round (PublicTree.py:244-251)
                c = new(kind, st.cur, a, s, i, me.depth + 1, me.cdepth)
            me.children.append(c)
        work.extend(reversed(pend))
    return nodes


def make_board_tables(rules, n_cdepth, root_board=()):
    """Boards per chance depth. boards[c] = int8 [nb_c, n_cards_out]; board_parent[c] = int32 [nb_c];
    children of one parent board are contiguous and ascend in (combination-)lexicographic card order."""
    deck = rules.N_CARDS_IN_DECK
    boards = [np.array([list(root_board)], dtype=np.int8).reshape(1, len(root_board))]
    parents = [np.zeros(1, np.int32)]
    first_round = {0: Poker.PREFLOP, rules.N_FLOP_CARDS: Poker.FLOP,
                   rules.N_FLOP_CARDS + rules.N_TURN_CARDS: Poker.TURN}.get(len(root_board), Poker.PREFLOP)
    rnd = first_round
    for c in range(1, n_cdepth + 1):
        rnd += 1
        k = rules.n_cards_dealt_in_transition_to(rnd)
        prev = boards[-1]
        rows, par = [], []
        for j in range(prev.shape[0]):
            used = set(prev[j].tolist())
            free = [x for x in range(deck) if x not in used]
            for combo in combinations(free, k):
                rows.append(list(prev[j]) + list(combo))
                par.append(j)
        boards.append(np.array(rows, dtype=np.int8).reshape(len(rows), prev.shape[1] + k))
        parents.append(np.array(par, dtype=np.int32))
    return boards, parents


class FlatTree:
    """Depth-sorted structure-of-arrays public tree (host numpy; uploaded to HBM by the solver)."""

    def __init__(self, game_cls, env_args, stop_at_street=None, board_tables=None, board_spec=None, root_actions=None):
        """board_spec (two-card games): a `holdem_boards.BoardSpec` (one chance layer: the boards dealt, their deal
        probability, their weight in the parent's sum, suit-permutation tables; default = all boards of the game as
        suit-isomorphism classes) or a `holdem_boards.MultiStreetBoards` (sub-game rooted at a fixed board with
        several chance layers).  root_actions: discrete actions played from the start of the hand to reach the root of a
        sub-game (the reference roots a tree at "the current state of the environment", PublicTree.py:37-40, 111-126)."""
        self.game_cls = game_cls
        self.rules = game_cls.RULES
        self.R = self.rules.RANGE_SIZE
        self.betting = eng.HUBetting(game_cls, env_args)
        root_state = None
        if root_actions:
            root_state = self.betting.reset()
            for a in root_actions:
                out, _ = self.betting.step(root_state, a)
                assert out in (eng.CONTINUE, eng.NEXT_ROUND), "root_actions must not end the hand"
        self.root_round = 0 if root_state is None else root_state.round
        self.abs_nodes = enumerate_betting_tree(self.betting, root_state=root_state, stop_at_street=stop_at_street)
        self.board_spec = None
        self.root_has_board = False
        if self.rules.N_HOLE_CARDS == 2:
            from pokerrl_b200.game.holdem_boards import BoardSpec, MultiStreetBoards
            n_cd = max(n.cdepth for n in self.abs_nodes)
            allin = [n for n in self.abs_nodes if n.kind == KIND_SHOWDOWN_ALLIN]
            if board_spec is None:
                if n_cd > 1 or root_state is not None:
                    raise ValueError("sub-games / multi-street two-card trees need an explicit MultiStreetBoards spec")
                board_spec = BoardSpec.full_game(self.rules)
            self.board_spec = board_spec
            # boards an all-in showdown before the deal runs out over (DeviceTree builds the equity matrix from them)
            self.allin_spec = board_spec if allin else None
            if n_cd == 0:  # e.g. a push / fold game: no chance node at all, the board spec only feeds the equity matrix
                assert isinstance(board_spec, BoardSpec)
                self.board_prob = np.ones(1, np.float32)
                self.board_mult = np.ones(1, np.float32)
                self._expand(([np.zeros((1, 0), np.int8)], [np.zeros(1, np.int32)]))
                return
            if isinstance(board_spec, MultiStreetBoards):
                assert board_spec.n_layers == n_cd, (board_spec.n_layers, n_cd)
                board_tables = (board_spec.boards, board_spec.parents)
                self.root_has_board = board_spec.boards[0].shape[1] > 0
                self.board_prob = np.concatenate(board_spec.prob).astype(np.float32)
                self.board_mult = np.concatenate(board_spec.mult).astype(np.float32)
            else:
                if n_cd != 1:
                    raise ValueError("a BoardSpec describes exactly one chance layer")
                board_tables = ([np.zeros((1, 0), np.int8), board_spec.boards],
                                [np.zeros(1, np.int32), np.zeros(board_spec.boards.shape[0], np.int32)])
                # per global board id (0 = the empty pre-deal board)
                self.board_prob = np.concatenate([[1.0], board_spec.board_prob]).astype(np.float32)
                self.board_mult = np.concatenate([[1.0], board_spec.board_mult]).astype(np.float32)
        self._expand(board_tables)

    # ------------------------------------------------------------------
    def _expand(self, board_tables):
        A = self.abs_nodes
        nA = len(A)
        a_kind = np.array([n.kind for n in A], np.int8)
        a_parent = np.array([n.parent for n in A], np.int64)
        a_depth = np.array([n.depth for n in A], np.int64)
        a_cdepth = np.array([n.cdepth for n in A], np.int64)
        a_nch = np.array([len(n.children) for n in A], np.int64)
        n_cd = int(a_cdepth.max())
        if board_tables is None:
            board_tables = make_board_tables(self.rules, n_cd)
        self.boards, self.board_parent = board_tables
        nb = np.array([b.shape[0] for b in self.boards], np.int64)
        fan = np.ones(n_cd + 1, np.int64)  # children boards per parent board, per chance depth
        for c in range(1, n_cd + 1):
            fan[c] = nb[c] // nb[c - 1]
            assert fan[c] * nb[c - 1] == nb[c]
        board_off = np.concatenate([[0], np.cumsum(nb)])  # global board id = board_off[c] + j
        self.board_off = board_off

        # sibling groups: one per abstract non-leaf parent, plus the root's own group
        a_k = np.zeros(nA, np.int64)  # index within sibling group
        a_m = np.ones(nA, np.int64)  # sibling group size
        for n in A:
            for k, c in enumerate(n.children):
                a_k[c] = k
                a_m[c] = len(n.children)
        # expanded subtree sizes (reference DFS numbering)
        size = np.ones(nA, np.int64)
        order = np.argsort(-a_depth, kind="stable")
        for i in order:
            n = A[i]
            if n.kind == KIND_CHANCE:
                size[i] = 1 + (fan[n.cdepth + 1] * size[n.children[0]] if n.children else 0)
            else:
                size[i] = 1 + sum(size[c] for c in n.children)
        a_reloff = np.zeros(nA, np.int64)  # DFS offset relative to parent (chance children: + j_local*size)
        for n in A:
            off = 1
            for c in n.children:
                a_reloff[c] = off
                off += size[c]

        # flat ids: levels = abstract depth; inside a level, groups in order of their parent abstract node
        max_depth = int(a_depth.max())
        a_base = np.zeros(nA, np.int64)  # base flat id of the sibling group of abstract node a
        level_start = [0]
        cursor = 0
        by_depth = [[] for _ in range(max_depth + 1)]
        for i, n in enumerate(A):
            by_depth[n.depth].append(i)
        for d in range(max_depth + 1):
            seen_parent = {}
            for i in by_depth[d]:  # DFS creation order keeps siblings adjacent & parents in order
                p = A[i].parent
                if p not in seen_parent:
                    seen_parent[p] = cursor
                    cursor += a_m[i] * nb[a_cdepth[i]]
                a_base[i] = seen_parent[p]
            level_start.append(cursor)
        N = cursor
        self.n_nodes = int(N)
        self.level_start = np.array(level_start, np.int64)
        self.n_levels = max_depth + 1

        parent = np.full(N, -1, np.int32)
        first_child = np.full(N, -1, np.int32)
        n_children = np.zeros(N, np.int32)
        kind = np.zeros(N, np.int8)
        acted_last = np.zeros(N, np.int8)
        action = np.full(N, -1, np.int32)
        pot = np.zeros(N, np.int64)
        rnd = np.zeros(N, np.int8)
        board = np.full(N, -1, np.int32)
        abs_id = np.zeros(N, np.int32)
        cdepth = np.zeros(N, np.int8)
        stack = np.zeros((N, 2), np.int64)
        bet = np.zeros((N, 2), np.int64)
        dfs_rel = np.zeros(N, np.int64)

        a_pot = np.array([n.pot for n in A], np.int64)
        a_round = np.array([n.round for n in A], np.int8)
        a_acted = np.array([n.acted_last for n in A], np.int8)
        a_action = np.array([n.action for n in A], np.int32)
        a_stack = np.array([n.stack for n in A], np.int64)
        a_bet = np.array([n.bet for n in A], np.int64)
        a_first_child_abs = np.array([n.children[0] if n.children else -1 for n in A], np.int64)

        for c in range(n_cd + 1):
            sel = np.nonzero(a_cdepth == c)[0]
            if sel.size == 0:
                continue
            J = np.arange(nb[c], dtype=np.int64)[None, :]
            ids = a_base[sel][:, None] + J * a_m[sel][:, None] + a_k[sel][:, None]  # [n_sel, nb_c]
            flat = ids.ravel()
            rep = lambda v: np.repeat(v[sel], nb[c])  # noqa: E731
            kind[flat] = rep(a_kind)
            acted_last[flat] = rep(a_acted)
            pot[flat] = rep(a_pot)
            rnd[flat] = rep(a_round)
            abs_id[flat] = np.repeat(sel.astype(np.int32), nb[c])
            cdepth[flat] = c
            stack[flat] = np.repeat(a_stack[sel], nb[c], axis=0)
            bet[flat] = np.repeat(a_bet[sel], nb[c], axis=0)
            board[flat] = (np.tile(board_off[c] + np.arange(nb[c]), sel.size)
                           if (c > 0 or getattr(self, "root_has_board", False)) else -1)
            # parents
            par_abs = a_parent[sel]
            has_par = par_abs >= 0
            pa = np.where(has_par, par_abs, 0)
            par_is_chance = has_par & (a_kind[pa] == KIND_CHANCE)
            j_par = np.where(par_is_chance[:, None], self.board_parent[c][None, :] if c > 0 else 0, J)
            pid = a_base[pa][:, None] + j_par * a_m[pa][:, None] + a_k[pa][:, None]
            pid = np.where(has_par[:, None], pid, -1)
            parent[flat] = pid.ravel()
            # action / DFS offset relative to parent
            j_local = (J - j_par * fan[c]) if c > 0 else np.zeros_like(J)
            act = np.where(par_is_chance[:, None], j_local, a_action[sel][:, None])
            action[flat] = act.ravel()
            rel = a_reloff[sel][:, None] + np.where(par_is_chance[:, None], j_local * size[sel][:, None], 0)
            dfs_rel[flat] = rel.ravel()
            # children
            fc_abs = a_first_child_abs[sel]
            has_ch = fc_abs >= 0
            fa = np.where(has_ch, fc_abs, 0)
            is_ch = a_kind[sel] == KIND_CHANCE
            nch = np.where(is_ch, fan[min(c + 1, n_cd)], a_nch[sel])
            fc = a_base[fa][:, None] + J * nch[:, None]
            first_child[flat] = np.where(has_ch[:, None], fc, -1).ravel()
            n_children[flat] = np.where(has_ch[:, None], np.broadcast_to(nch[:, None], ids.shape), 0).ravel()

        # absolute DFS index, top-down by level
        dfs = np.zeros(N, np.int64)
        for d in range(1, self.n_levels):
            lo, hi = self.level_start[d], self.level_start[d + 1]
            dfs[lo:hi] = dfs[parent[lo:hi]] + dfs_rel[lo:hi]
        # table slots: children of decision nodes, in flat order (so one decision node's rows are contiguous)
        is_dec_child = np.zeros(N, bool)
        nz = parent >= 0
        is_dec_child[nz] = kind[parent[nz]] <= KIND_P1
        slot = np.full(N, -1, np.int32)
        slot[is_dec_child] = np.arange(int(is_dec_child.sum()), dtype=np.int32)
        first_slot = np.full(N, -1, np.int32)
        dec = (kind <= KIND_P1) & (first_child >= 0)
        first_slot[dec] = slot[first_child[dec]]

        self.parent, self.first_child, self.n_children = parent, first_child, n_children
        self.kind, self.acted_last, self.action = kind, acted_last, action
        self.pot, self.round, self.board, self.abs_id = pot, rnd, board, abs_id
        self.cdepth = cdepth  # number of chance deals above the node
        self._abs_base, self._abs_m, self._abs_k = a_base, a_m, a_k  # flat id of (abstract node a, board j) = base + j*m + k
        self.stack, self.bet = stack, bet
        self.dfs = dfs
        self.slot, self.first_slot = slot, first_slot
        self.n_slots = int(is_dec_child.sum())
        self.n_nonterm = int(((kind <= KIND_CHANCE) & (first_child >= 0)).sum())
        self.n_decision = int(dec.sum())
        self.max_actions = int(n_children[dec].max()) if dec.any() else 0

    # ------------------------------------------------------------------ helpers
    def board_cards(self):
        """int8 [n_boards_total, N_TOTAL_BOARD_CARDS] padded with the not-dealt token, global board id order."""
        nmax = max(self.rules.N_TOTAL_BOARD_CARDS, 1)
        rows = []
        for b in self.boards:
            pad = np.full((b.shape[0], nmax), Poker.CARD_NOT_DEALT_TOKEN_1D, np.int8)
            pad[:, :b.shape[1]] = b
            rows.append(pad)
        return np.concatenate(rows, axis=0)

    def node_board_cards(self):
        """int8 [N, N_TOTAL_BOARD_CARDS]; not-dealt token where no board."""
        bc = self.board_cards()
        out = np.full((self.n_nodes, bc.shape[1]), Poker.CARD_NOT_DEALT_TOKEN_1D, np.int8)
        m = self.board >= 0
        # board == -1 means chance depth 0 -> global board 0 (empty)
        out[m] = bc[self.board[m]]
        return out

    def allin_completions(self):
        """Two-card games: for every public board an all-in showdown happens on BEFORE the board is complete, the complete
        boards it runs out over and their weights (product of deal probability x weight in the parent's sum over the missing
        deals) - the inputs of that board's equity matrix (csrc/allin_dense.cu; one-card analogue ValueFiller.py:160-175).
        Returns {board key: (boards int8 [n, 5], weights float64 [n], sym_perm or None)}; key = the nodes' `board` entry
        (-1: no board yet)."""
        from pokerrl_b200.game.holdem_boards import MultiStreetBoards
        nodes = np.nonzero(self.kind == KIND_SHOWDOWN_ALLIN)[0]
        out = {}
        if nodes.size == 0:
            return out
        spec = self.allin_spec
        if not isinstance(spec, MultiStreetBoards):  # one deal: every board of the spec
            w = np.asarray(spec.board_prob, np.float64) * np.asarray(spec.board_mult, np.float64)
            out[int(self.board[nodes[0]])] = (np.ascontiguousarray(spec.boards, np.int8), w, spec.sym_perm)
            assert np.all(self.board[nodes] == self.board[nodes[0]])
            return out
        L = spec.n_layers
        for key in np.unique(self.board[nodes]):
            gid = max(int(key), 0)  # -1 = the root's (empty) board = global id 0
            c = int(np.searchsorted(self.board_off, gid, side="right") - 1)  # layer of that board
            idx, w = np.array([gid - int(self.board_off[c])], np.int64), np.ones(1)
            for layer in range(c + 1, L + 1):  # children of the current set, layer by layer
                par = np.asarray(spec.parents[layer], np.int64)
                pos = {int(j): k for k, j in enumerate(idx)}
                sel = np.nonzero(np.isin(par, idx))[0]
                wl = np.asarray(spec.prob[layer], np.float64)[sel] * np.asarray(spec.mult[layer], np.float64)[sel]
                w = w[[pos[int(j)] for j in par[sel]]] * wl
                idx = sel
            assert spec.boards[L].shape[1] == self.rules.N_TOTAL_BOARD_CARDS, "the last layer must complete the board"
            out[int(key)] = (np.ascontiguousarray(spec.boards[L][idx], np.int8), w, None)
        return out

    def board_subtree(self):
        """Description of the post-deal subtree shared by all boards of a single-chance-layer tree (prl_subtree_t):
        local nodes = abstract nodes below the deal in breadth-first order.  None if the tree has another shape."""
        A = self.abs_nodes
        ch = [i for i, n in enumerate(A) if n.kind == KIND_CHANCE]
        if len(ch) != 1 or max(n.cdepth for n in A) != 1:
            return None
        local = sorted([i for i, n in enumerate(A) if n.cdepth == 1], key=lambda i: (A[i].depth, i))
        if len(local) > 16:
            return None
        loc = {a: i for i, a in enumerate(local)}
        chance_flat = int(self._abs_base[ch[0]] + self._abs_k[ch[0]])
        return dict(
            n_local=len(local), chance_node=chance_flat, chance_level=int(A[ch[0]].depth),
            n_boards_local=int(self.n_children[chance_flat]), first_board=int(self.board[self.first_child[chance_flat]]),
            node_base=[int(self._abs_base[a]) for a in local], node_m=[int(self._abs_m[a]) for a in local],
            node_k=[int(self._abs_k[a]) for a in local], kind=[int(A[a].kind) for a in local],
            parent=[loc.get(A[a].parent, -1) for a in local],
            first_child=[loc[A[a].children[0]] if A[a].children else -1 for a in local],
            n_children=[len(A[a].children) for a in local], acted_last=[int(A[a].acted_last) for a in local],
            pot=[float(A[a].pot) for a in local])

    def work_order(self):
        """(order, level_nonterm): per level the node ids sorted by (kind, n_children), non-terminals first, so that
        a warp of the GPU sweeps holds nodes of one kind; level_nonterm[d] = number of non-terminals of level d."""
        order = np.empty(self.n_nodes, np.int32)
        level_nonterm = np.zeros(self.n_levels, np.int64)
        # (kind, chance depth, fan-out): among the chance nodes of a level those closest to the root come first - they
        # are the ones whose children are sharded over GPUs (pokerrl_b200/distributed.py)
        key = (self.kind.astype(np.int64) * (1 << 40) + self.cdepth.astype(np.int64) * (1 << 32)
               + self.n_children.astype(np.int64))
        for d in range(self.n_levels):
            lo, hi = int(self.level_start[d]), int(self.level_start[d + 1])
            order[lo:hi] = lo + np.argsort(key[lo:hi], kind="stable")
            level_nonterm[d] = int((self.kind[lo:hi] <= KIND_CHANCE).sum())
        return order, level_nonterm

    def dfs_permutation(self):
        """perm such that array_in_dfs_order = array_in_flat_order[perm]"""
        perm = np.empty(self.n_nodes, np.int64)
        perm[self.dfs] = np.arange(self.n_nodes)
        return perm
