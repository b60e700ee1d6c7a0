"""Minimal stand-ins for gymnasium.spaces used when gymnasium is not installed (the reference
builds its spaces with gymnasium, env_hetero.py:29-43).  If gymnasium is importable the real
classes are used so RLlib-style drivers see genuine spaces."""
import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium.spaces import Box, Dict, MultiDiscrete, Discrete  # type: ignore
except Exception:  # noqa: BLE001
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

    class MultiDiscrete:
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, np.int64)
            self.shape = self.nvec.shape
            self.dtype = np.dtype(np.int64)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def sample(self):
            return (np.random.random(self.shape) * self.nvec).astype(np.int64)

    class Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()

        def contains(self, x):
            return 0 <= int(x) < self.n

        def sample(self):
            return int(np.random.randint(self.n))

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def contains(self, x):
            return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}
