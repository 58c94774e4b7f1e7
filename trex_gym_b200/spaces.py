"""Minimal ``gym.spaces.Box`` / ``gym.Env`` stand-ins, used only when ``gym`` is not installed
(it is not in this image).  With gym present the real classes are used."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    import gym as _gym
    from gym import spaces as _spaces

    Env = _gym.Env
    Box = _spaces.Box
    HAVE_GYM = True
except Exception:  # noqa: BLE001
    HAVE_GYM = False

    class Env(object):
        metadata = {}
        reward_range = (-float("inf"), float("inf"))
        action_space = None
        observation_space = None

        def step(self, action):
            raise NotImplementedError

        def reset(self):
            raise NotImplementedError

        def render(self, mode="human"):
            raise NotImplementedError

        def close(self):
            pass

        def seed(self, seed=None):
            return [seed]

    class Box(object):
        """``Box(low, high, dtype)`` with array bounds (the only form trex_env.py:94-96 uses)."""

        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            low = np.asarray(low, dtype=self.dtype)
            high = np.asarray(high, dtype=self.dtype)
            if shape is not None and low.shape == ():
                low = np.full(shape, low, dtype=self.dtype)
                high = np.full(shape, high, dtype=self.dtype)
            assert low.shape == high.shape
            self.low, self.high = low, high
            self.shape = low.shape
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return "Box%s" % (self.shape,)


def np_random(seed=None):
    """``gym.utils.seeding.np_random`` (trex_env.py:125)."""
    if HAVE_GYM:  # pragma: no cover
        from gym.utils import seeding

        return seeding.np_random(seed)
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1)[0])
    return np.random.RandomState(seed % (2 ** 32)), seed
