"""Observation/action space objects for the MultiAgentEnv surface (core.py:117-148).

Uses `gym.spaces` when gym is installed (as RLlib expects); otherwise a small self-contained
implementation with the same `contains` / `sample` / `shape` / `dtype` / `n` behaviour, so that
`baseline/run.py`'s `observation_space.contains(obs)` asserts (run.py:36-37,58-59) work without gym.
"""
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    from gym.spaces import Box, Dict, Discrete, MultiBinary, Space  # type: ignore  # noqa: F401

    HAVE_GYM = True
except Exception:  # noqa: BLE001
    HAVE_GYM = False
    _rng = np.random.RandomState()

    class Space:
        shape = None
        dtype = None

        def contains(self, x):
            raise NotImplementedError

        def sample(self):
            raise NotImplementedError

        def __contains__(self, x):
            return self.contains(x)

    class Discrete(Space):
        def __init__(self, n):
            self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)

        def contains(self, x):
            if isinstance(x, (int, np.integer)) or (isinstance(x, np.ndarray) and x.shape == () and x.dtype.kind in "iu"):
                return 0 <= int(x) < self.n
            return False

        def sample(self):
            return int(_rng.randint(self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
                        and np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            return _rng.randint(self.low, self.high + 1).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class MultiBinary(Space):
        def __init__(self, n):
            self.n, self.shape, self.dtype = int(n), (int(n),), np.dtype(np.int8)

        def contains(self, x):
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))

        def sample(self):
            return _rng.randint(0, 2, size=self.n).astype(np.int8)

        def __repr__(self):
            return f"MultiBinary({self.n})"

    class Dict(Space):
        def __init__(self, spaces):
            self.spaces = OrderedDict(sorted(spaces.items()))

        def contains(self, x):
            if not isinstance(x, dict) or len(x) != len(self.spaces):
                return False
            return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def sample(self):
            return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

        def __getitem__(self, k):
            return self.spaces[k]

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k}:{s!r}" for k, s in self.spaces.items()) + ")"


def observation_space(num_requests: int, area_dimension: int):
    """core.py:119-148."""
    R, dim, i32 = num_requests, area_dimension, np.int32
    return Dict({
        "num_agents": Box(low=1, high=R, shape=(1,), dtype=i32),
        "self_position": Box(low=0, high=dim, shape=(2,), dtype=i32),
        "self_availability": MultiBinary(1),
        "self_delivery_target": Box(low=0, high=dim, shape=(2,), dtype=i32),
        "other_positions": Box(low=0, high=dim, shape=(R - 1, 2), dtype=i32),
        "other_availabilities": MultiBinary(R - 1),
        "other_delivery_targets": Box(low=0, high=dim, shape=(R - 1, 2), dtype=i32),
        "requests": Box(low=0, high=dim, shape=(R, 4), dtype=i32),
    })
