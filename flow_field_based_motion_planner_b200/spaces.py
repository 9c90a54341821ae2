"""Minimal Box / Dict spaces with the attribute surface of the 2020 `gym.spaces` the reference used
(/root/reference/src/gym_ffmp/envs/ffmp.py:32,55-64).  `gym` itself is not a dependency."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        low = np.asarray(low)
        high = np.asarray(high)
        if shape is not None:
            low = np.broadcast_to(low, shape)
            high = np.broadcast_to(high, shape)
        self.low = low.astype(self.dtype)
        self.high = high.astype(self.dtype)
        self.shape = tuple(self.low.shape)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        return rng.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x) -> bool:
        return 0 <= int(x) < self.n

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        return int(rng.integers(0, self.n))

    def __repr__(self):
        return f"Discrete({self.n})"


class Dict:
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, key):
        return self.spaces[key]

    def keys(self):
        return self.spaces.keys()

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k}: {v!r}" for k, v in self.spaces.items()) + ")"
