#!/usr/bin/env python
"""Latency of the single-env class (the reference's Game2048Env API on the GPU): one g2048_one call = one kernel
launch + one stream synchronisation per method.  Compared with the unmodified reference on one host core when
baseline/_ref is installed.   python scripts/single_env_latency.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gym_2048_b200 as g  # noqa: E402


def play(env, n, seed=0):
    rng = np.random.default_rng(seed)
    acts = rng.integers(0, 4, n)
    env.reset(seed=seed)
    t0 = time.perf_counter()
    for a in acts:
        _, _, term, _, _ = env.step(int(a))
        if term:
            env.reset()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    e = g.Game2048Env()
    play(e, 200)
    print("gym_2048_b200.Game2048Env.step() (+ reset on done): %.1f us per call" % play(e, 5000))
    t0 = time.perf_counter()
    for _ in range(2000):
        e.highest()
    print("highest(): %.1f us per call" % ((time.perf_counter() - t0) / 2000 * 1e6))
    t0 = time.perf_counter()
    for _ in range(2000):
        e.move(0, trial=True) if e.legal_actions() & 1 else None
    print("legal_actions() [+ move(0, trial=True) when legal]: %.1f us per iteration" % ((time.perf_counter() - t0) / 2000 * 1e6))
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "env")):
        try:
            import gymnasium  # noqa: F401
        except ImportError:
            sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
        sys.path.insert(0, ref)
        for k in [k for k in sys.modules if k == "env" or k.startswith("env.")]:
            del sys.modules[k]
        from env.envs.game2048_env import Game2048Env as RefEnv
        r = RefEnv()
        play(r, 200)
        print("reference Game2048Env.step() on one host core: %.1f us per call" % play(r, 5000))


if __name__ == "__main__":
    main()
