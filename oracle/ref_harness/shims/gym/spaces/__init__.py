import numpy as np


class Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            shape = low.shape
        else:
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return self.np_random.uniform(low=self.low, high=self.high,
                                      size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)
