"""gym.spaces stand-in (see package docstring). Semantics follow gym 0.21."""
from collections import OrderedDict

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)

    def contains(self, x):
        raise NotImplementedError

    def sample(self):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        # gym seeds its own RandomState per space; the global stream is NOT used.
        return int(_space_rng.randint(self.n))

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            v = int(x)
        elif isinstance(x, np.ndarray) and x.dtype.kind in "iu" and x.shape == ():
            v = int(x)
        else:
            return False
        return 0 <= v < self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(shape, dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(
            x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
            and np.all(x >= self.low) and np.all(x <= self.high)
        )

    def sample(self):
        return _space_rng.randint(self.low, self.high + 1).astype(self.dtype)


class MultiBinary(Space):
    def __init__(self, n):
        super().__init__((int(n),), np.int8)
        self.n = int(n)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))

    def sample(self):
        return _space_rng.randint(0, 2, size=self.n).astype(np.int8)


class Dict(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = OrderedDict(sorted(spaces.items()))

    def contains(self, x):
        if not isinstance(x, dict) or len(x) != len(self.spaces):
            return False
        return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())


_space_rng = np.random.RandomState(0)
