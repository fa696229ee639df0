"""Minimal gym.spaces look-alikes (gym 0.10.5 is not installed offline; the reference uses
spaces.Discrete and spaces.Box: ssf_env.py:90,169, rl/envs.py:21-25). If the real `gym` package is
importable its classes are used instead, so isinstance checks in caller code keep working."""
import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    from gym.spaces import Box, Discrete  # type: ignore
except Exception:
    class Discrete(object):
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return "Discrete(%d)" % self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and other.n == self.n

    class Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, self.dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high, self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def __repr__(self):
            return "Box%s" % (self.shape,)
