"""BASELINE config 1 (plumbing, no GPU): the UNMODIFIED reference env, imported through the gymnasium stand-in
of oracle/shim, plays a 1000-step random-action rollout with its own RNG (numpy PCG64).  The observed seeded
boards and the SHA-256 of the rollout are the values SURVEY.md §8c recorded from the same harness, so this pins
the harness every golden fixture was generated with (stand-in seeding included).  Skipped where neither
/root/reference nor baseline/_ref holds the reference package."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_env_class():
    for base in (os.environ.get("G2048_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(base, "env", "envs", "game2048_env.py")):
            try:
                import gymnasium  # noqa: F401
            except ImportError:
                sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
            sys.path.insert(0, base)
            for name in [m for m in sys.modules if m == "env" or m.startswith("env.")]:
                del sys.modules[name]
            from env.envs.game2048_env import Game2048Env
            return Game2048Env
    return None


def test_config1_reference_rollout_hash_and_seeded_boards():
    Env = _reference_env_class()
    if Env is None:
        pytest.skip("reference package not present")
    for seed, cells in ((0, [2, 14]), (1, [0, 1]), (42, [6, 13]), (456, [1, 6])):
        e = Env()
        e.reset(seed=seed)
        flat = np.asarray(e.get_board()).reshape(-1)
        assert list(np.flatnonzero(flat)) == cells and set(flat[flat != 0]) == {2}
    env = Env()
    env.reset(seed=0)
    h = hashlib.sha256()
    dones = 0
    for a in np.random.default_rng(123).integers(0, 4, 1000):
        obs, reward, terminated, truncated, info = env.step(int(a))
        assert isinstance(reward, float) and isinstance(terminated, bool) and truncated is False
        assert obs.shape == (16, 4, 4) and int(obs.sum()) == 16 - int((np.asarray(env.get_board()) >= 65536).sum())
        h.update(np.asarray(env.get_board()).astype(np.int64).tobytes())
        h.update(np.float64(reward).tobytes())
        h.update(bytes([1 if terminated else 0]))
        if terminated:
            dones += 1
            env.reset()
    assert dones == 72
    assert h.hexdigest() == "5e67d26991e0d1e7432469f6e6f44ede0ab54271bef2e00b28cb65677a28d5e2"
