import numpy as np


class Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.low = np.asarray(low).astype(self.dtype)
        self.high = np.asarray(high).astype(self.dtype)
        self.shape = self.low.shape


class Dict(object):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]
