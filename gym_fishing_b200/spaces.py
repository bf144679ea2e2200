"""Observation/action space descriptors with the attribute surface the reference's callers touch
(`.low .high .dtype .shape .n .sample() .contains()`; gym_fishing/envs/base_fishing_env.py:49-58,
fishing_env.py:24, shared_env.py:71,84-88).  If the real `gym` or `gymnasium` is importable its classes are used,
so isinstance checks in third-party code (stable-baselines3) keep working; otherwise these stand-alone ones."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - neither package ships in the build image
    from gymnasium.spaces import Box, Discrete  # type: ignore
    BACKEND = "gymnasium"
except Exception:  # noqa: BLE001
    try:
        from gym.spaces import Box, Discrete  # type: ignore
        BACKEND = "gym"
    except Exception:  # noqa: BLE001
        BACKEND = "builtin"

        class Box:  # type: ignore[no-redef]
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                self.low = np.asarray(low, dtype=self.dtype)
                self.high = np.asarray(high, dtype=self.dtype)
                if shape is not None:
                    self.low = np.broadcast_to(self.low, shape).copy()
                    self.high = np.broadcast_to(self.high, shape).copy()
                self.shape = self.low.shape
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            __contains__ = contains

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

            def __eq__(self, other):
                return (isinstance(other, Box) and self.shape == other.shape and np.array_equal(self.low, other.low)
                        and np.array_equal(self.high, other.high))

        class Discrete:  # type: ignore[no-redef]
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.dtype(np.int64)
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                return int(self._rng.integers(self.n))

            def contains(self, x):
                try:
                    return 0 <= int(x) < self.n and int(x) == x
                except (TypeError, ValueError):
                    return False

            __contains__ = contains

            def __repr__(self):
                return f"Discrete({self.n})"

            def __eq__(self, other):
                return isinstance(other, Discrete) and self.n == other.n


def unit_box():
    """The [-1, 1] float32 box used for every observation space and every continuous action space
    (base_fishing_env.py:49-58)."""
    return Box(np.array([-1], dtype=np.float32), np.array([1], dtype=np.float32), dtype=np.float32)
