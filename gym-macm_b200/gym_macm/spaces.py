"""`spaces` for the env hosts: gym's or gymnasium's when one is importable, otherwise a minimal
stand-in with the three methods the reference relies on (`contains`, `sample`, item access).
The reference builds Dict/MultiDiscrete/Box/Tuple/Discrete spaces in create_space()
(mvmnt.py:142-158, combat.py:186-201) and only ever calls `action_space.contains(actions)`
(mvmnt.py:94, combat.py:118)."""
import numpy as np

try:  # pragma: no cover - neither package exists in the build container
    from gym import spaces as _sp
    Dict, MultiDiscrete, Box, Tuple, Discrete = _sp.Dict, _sp.MultiDiscrete, _sp.Box, _sp.Tuple, _sp.Discrete
    BACKEND = "gym"
except Exception:
    try:  # pragma: no cover
        from gymnasium import spaces as _sp
        Dict, MultiDiscrete, Box, Tuple, Discrete = _sp.Dict, _sp.MultiDiscrete, _sp.Box, _sp.Tuple, _sp.Discrete
        BACKEND = "gymnasium"
    except Exception:
        BACKEND = "builtin"

        class Discrete(object):
            def __init__(self, n):
                self.n = int(n)

            def contains(self, x):
                try:
                    return int(x) == x and 0 <= int(x) < self.n
                except Exception:
                    return False

            def sample(self):
                return int(np.random.randint(self.n))

        class MultiDiscrete(object):
            def __init__(self, nvec):
                self.nvec = np.asarray(nvec, dtype=np.int64)

            def contains(self, x):
                x = np.asarray(x)
                return (x.shape == self.nvec.shape and np.issubdtype(x.dtype, np.integer)
                        and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec)))

            def sample(self):
                return (np.random.random_sample(self.nvec.shape) * self.nvec).astype(np.int64)

        class Box(object):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape if shape is not None else np.shape(low))
                self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.low.shape)
                self.shape = self.low.shape

            def contains(self, x):
                x = np.asarray(x, dtype=np.float64)
                return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1e3)
                hi = np.where(np.isfinite(self.high), self.high, 1e3)
                return np.random.uniform(lo, hi)

        class Tuple(object):
            def __init__(self, spaces):
                self.spaces = tuple(spaces)

            def contains(self, x):
                return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

            def sample(self):
                return tuple(s.sample() for s in self.spaces)

        class Dict(object):
            def __init__(self, spaces):
                self.spaces = dict(spaces)

            def __getitem__(self, k):
                return self.spaces[k]

            def keys(self):
                return self.spaces.keys()

            def contains(self, x):
                if not isinstance(x, dict) or len(x) != len(self.spaces):
                    return False
                return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

            def sample(self):
                return {k: s.sample() for k, s in self.spaces.items()}
