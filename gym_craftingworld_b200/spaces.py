"""Minimal stand-ins for the ``gym.spaces`` metadata the reference exposes (``ray.py:84-110, 133``).

``gym`` is not a dependency of this package (it is not installable in the target image); these classes carry the
same attributes callers read (``n``, ``shape``, ``low``, ``high``, ``dtype``, ``spaces``, ``sample()``).
"""
from __future__ import annotations

import numpy as np


class Discrete:
    def __init__(self, n: int):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.RandomState()

    def sample(self):
        return int(self._rng.randint(self.n))

    def contains(self, x) -> bool:
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class Box:
    def __init__(self, low, high, shape, dtype=np.uint8):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.low, self.high = low, high          # scalars: bounds are uniform (0..255 or 0..1 upstream)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class Dict:
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k}: {v}" for k, v in self.spaces.items()) + ")"
