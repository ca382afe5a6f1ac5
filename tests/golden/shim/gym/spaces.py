import numpy as np


class Discrete(object):
    def __init__(self, n):
        self.n = n

    def contains(self, x):
        return 0 <= int(x) < self.n


class MultiDiscrete(object):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))


class Box(object):
    def __init__(self, low=None, high=None, shape=None):
        self.low, self.high = np.asarray(low, dtype=float), np.asarray(high, dtype=float)

    def contains(self, x):
        x = np.asarray(x, dtype=float)
        return bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))


class Tuple(object):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)


class Dict(object):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def contains(self, x):
        return isinstance(x, dict) and set(x) == set(self.spaces) and all(self.spaces[k].contains(v) for k, v in x.items())
