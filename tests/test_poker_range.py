"""`pokerrl_b200.game.PokerRange` against a trace of the REFERENCE's PokerRange (tests/golden/poker_range.npz, produced by
oracle/gen_golden_poker_range.py): reset, blocker removal, multiplications with renormalisation (incl. the fall-back to a
uniform range when everything is multiplied away), street changes - float32 ranges bit for bit.  CPU only."""
import os

import numpy as np

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poker_range.npz"))


def _vector(seed, n, cols=None):
    r = np.random.default_rng(seed)
    return (r.random(n if cols is None else (n, cols)) ** 2).astype(np.float32)


def test_poker_range_replays_the_reference_trace():
    from pokerrl_b200.game import bet_sets, games
    from pokerrl_b200.game.PokerRange import PokerRange
    from pokerrl_b200.game.wrappers import VanillaEnvBuilder
    g = games.DiscretizedNLHoldem
    bldr = VanillaEnvBuilder(g, g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[20000, 20000], bet_sizes_list_as_frac_of_pot=bet_sets.B_2))
    lut = bldr.lut_holder
    r = PokerRange(env_bldr=bldr)
    for k in range(len(GOLD["op"])):
        op, a, b, cards = int(GOLD["op"][k]), int(GOLD["a"][k]), int(GOLD["b"][k]), GOLD["cards"][k]
        if op == 0:
            r.reset()
        elif op == 1:
            r.set_cards_to_zero_prob(cards_2d=lut.get_2d_cards(cards[cards >= 0]))
        elif op == 2:
            r.mul_and_norm(np.zeros(1326, np.float32) if a == 1 else _vector(b, 1326))
        elif op == 3:
            r.update_after_action(action=a, all_a_probs_for_all_hands=_vector(b, 1326, 3))
        else:
            n_out = {1: 3, 2: 4, 3: 5}[a]
            bd = np.full(5, -127, np.int8)
            bd[:n_out] = cards[:n_out]
            r.update_after_new_round(new_round=a, board_now_2d=lut.get_2d_cards(bd))
        assert r.range.dtype == np.float32 and np.array_equal(r.range, GOLD["ranges"][k]), (k, op)
    assert abs(float(r.get_card_probs().sum()) - 2.0) < 1e-5  # every hand holds two cards
