"""Writes pokerrl_b200/game/data/flop5_iso_classes.npz: the suit-isomorphism classes of all C(52,5) boards (representatives in
lexicographic order + orbit sizes), computed by holdem_boards.canonical_boards.  Loaded by BoardSpec.full_game."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pokerrl_b200.game.games import FlopHoldemRules  # noqa: E402
from pokerrl_b200.game.holdem_boards import _combos_52_5, canonical_boards  # noqa: E402

reps, orbit = canonical_boards(_combos_52_5(), FlopHoldemRules.N_SUITS)
assert reps.shape == (134459, 5) and int(orbit.sum()) == 2598960
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pokerrl_b200", "game", "data")
os.makedirs(out, exist_ok=True)
np.savez_compressed(os.path.join(out, "flop5_iso_classes.npz"), boards=reps, orbit=orbit.astype(np.int8))
print("wrote", reps.shape[0], "classes")
