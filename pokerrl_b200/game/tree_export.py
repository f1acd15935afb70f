"""PokerViz tree export: the nested dict `PublicTree.get_tree_as_dict()` produces
(PokerRL/game/_/tree/PublicTree.py:143-149, 302-420) - same keys, same strings, same float formatting - from the flat
tree and host copies of the device vectors.  Wire format only; nothing here touches the compute path.

Quirks of the reference kept on purpose (the golden files in tests/golden/export_*.json.gz come from the reference):
  * 'terminal' is always "TERM " + str(allowed_actions)  (`"TERM" if str(node.is_terminal)` is always true, :385)
  * 'data' is always "DATA: "  (the isinstance test at :370 looks at the node, not at node.data)
  * per-node exploitability is printed with numpy's str() of a float32 pair; "BR Action per hand" is filled at decision
    nodes only (chance and terminal nodes keep br_a_idx None, ValueFiller.py:34-93)
"""
import json

import numpy as np

from pokerrl_b200.game.Poker import Poker

KIND_P1, KIND_CHANCE, KIND_FOLD = 1, 2, 3

_HOLDEM_RANKS = ["2", "3", "4", "5", "6", "7", "8", "9", "T", "J", "Q", "K", "A"]
_HOLDEM_SUITS = ["h", "d", "s", "c"]


def card_str(rules, c):
    """game_rules.py RANK_DICT / SUIT_DICT of the rule set (:45-48, 110-113, 182-204), 1D card id -> e.g. "2a", "Th" """
    r, s = int(c) // rules.N_SUITS, int(c) % rules.N_SUITS
    if rules.N_HOLE_CARDS == 2:
        return _HOLDEM_RANKS[r] + _HOLDEM_SUITS[s]
    if rules.N_SUITS < 8:
        return str(r + 2) + "abcdefg"[s]
    return str(r + 2) + ("_" if rules.STRING == "BIG_LEDUC_RULES" else "") + str(s)


def cards2str(rules, cards_1d, seperator=", "):
    """PokerEnv.cards2str (PokerEnv.py:1294-1311) for 1D card ids; undealt slots are skipped"""
    return "".join(card_str(rules, c) + seperator for c in cards_1d if c >= 0)


def action_str(action):
    """PublicTree._get_action_as_str (:302-311): RX = X-th bet size of the node's bet set"""
    if action is None:
        return "None"
    if isinstance(action, str):
        return action  # "CHANCE"
    if action == Poker.FOLD:
        return "FOLD"
    if action == Poker.CHECK_CALL:
        return "CHECK"
    return "R" + str(action - 2)


def _arr2str(arr):
    if arr is None:
        return "Not Computed"
    out = ""
    for i in range(arr.shape[0]):
        for j in range(arr.shape[1]):
            out += str("{:10.4f}".format(arr[i, j])) + " "
        if i < arr.shape[0] - 1:
            out += " || "
    return out


def export_tree_dict(ft, reach=None, ev=None, ev_br=None, strategy_of=None):
    """ft: FlatTree (heads-up).  reach / ev / ev_br: float32 [2, n_nodes, >= R] in flat node order or None (not computed
    yet); strategy_of(n) -> [R, A] array of the node's acting player (chance included) or None."""
    rules, R = ft.rules, ft.R
    boards = ft.node_board_cards()
    valued = ev is not None and ev_br is not None and reach is not None

    def rec(n):
        kind = int(ft.kind[n])
        fc, A = int(ft.first_child[n]), int(ft.n_children[n])
        decision = kind <= KIND_P1 and fc >= 0
        allowed = [int(a) for a in ft.action[fc:fc + A]] if decision else []
        acted = int(ft.acted_last[n])
        acted_last = None if acted == -2 else ("Ch" if acted == -1 else acted)
        nxt = kind if kind <= KIND_P1 else ("Ch" if kind == KIND_CHANCE else None)
        board = cards2str(rules, boards[n])
        if n == 0:
            title = "ROOT"
        elif acted == -1:  # ChanceNode: the node right after a deal
            title = board
        else:
            title = "Player acted last " + str(acted_last) + " :: Action: " + action_str(int(ft.action[n])) + \
                    " :: Board: " + board
        playing = [1, 1]
        if kind == KIND_FOLD:
            playing[acted] = 0
        if valued:
            r_, e_, b_ = reach[:, n, :R], ev[:, n, :R], ev_br[:, n, :R]
            expl = str(np.sum(b_ * r_ - e_ * r_, axis=1))
            br = ""
            if decision:
                best = np.argmax(ev_br[kind, fc:fc + A, :R], axis=0)
                br = str([action_str(a) for a in np.array(allowed)[best]])
        else:
            r_ = None if reach is None else reach[:, n, :R]
            if reach is None and n == 0:  # build_tree seeds the root with the uniform prior (PublicTree.py:122-124)
                r_ = np.full((2, R), 1.0 / float(R), dtype=np.float32)
            e_ = b_ = None
            expl, br = "Exploitability not computed", ""
        strat = strategy_of(n) if (strategy_of is not None and kind <= KIND_CHANCE and fc >= 0) else None
        return {
            "text": {
                "title": title,
                "round": "Round : " + Poker.INT2STRING_ROUND[int(ft.round[n])],
                "main_pot": "Pot : " + json.dumps(int(ft.pot[n])),
                "terminal": "TERM " + str(allowed),
                "side_pots": "SP: " + json.dumps([0] * 2),
                "stack_sizes": "Stacks: " + json.dumps([int(s) for s in ft.stack[n]]),
                "current_bets": "Bets: " + json.dumps([int(b) for b in ft.bet[n]]),
                "not_folded": "Playing: " + json.dumps(playing) + "  Next: " + str(nxt),
                "exploitability": "Exploitability: " + expl + "   ||   BR Action per hand " + br,
                "strategy": "STRAT: " + _arr2str(strat),
                "reach_probs": "REACH: " + _arr2str(r_),
                "ev": "EV: " + _arr2str(e_),
                "ev_br": "EV-BR: " + _arr2str(b_),
                "data": "DATA: ",
            },
            "collapsed": True,
            "children": [rec(c) for c in range(fc, fc + A)] if fc >= 0 else [],
        }

    return rec(0)


def write_tree_js(path, dictionary):
    """file_util.write_dict_to_file_js (PokerRL/util/file_util.py:36-39): the `data.js` PokerViz loads"""
    with open(path, "w") as f:
        f.write("const data=" + json.dumps(dictionary))
