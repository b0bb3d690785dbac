"""Minimal stand-in for the `gymnasium` package — TEST INFRASTRUCTURE ONLY.

gymnasium is not installed in this image and there is no network.  The unmodified
reference env (`/root/reference/env/envs/game2048_env.py:3-4`, `env/__init__.py:1`)
needs only: `gymnasium.Env` (with the seeding `reset(seed)` that installs
`np_random = numpy.random.Generator(PCG64(SeedSequence(seed)))`, an assignable
`np_random`, `.unwrapped`, `.close()`), `gymnasium.spaces.Discrete/Box`,
`gymnasium.envs.registration.register` and `gymnasium.make`.  This module provides
exactly that so the reference can be imported as the oracle of record and timed as
the CPU baseline.  It is put on `sys.path` only by tests/golden/make_golden.py,
tests/ and bench.py's reference arm, and only when the real gymnasium is absent.
"""
import importlib

import numpy as np

from . import spaces  # noqa: F401
from .envs import registration  # noqa: F401
from .envs.registration import make, register  # noqa: F401

__version__ = "0.0-shim"


class Env:
    metadata = {}
    render_mode = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    def step(self, action):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self):
        pass
