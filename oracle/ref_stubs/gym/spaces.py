"""`gym.spaces` stub: just enough surface for PokerRL/game/_/rl_env/base/PokerEnv.py."""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else ()


class Tuple:
    def __init__(self, spaces):
        self.spaces = tuple(spaces)
        self.shape = None
