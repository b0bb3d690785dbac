#!/usr/bin/env python
"""e2e (host-buffer) step time of HostSteppedEnv for chunk counts and board formats.  G2048_SO=<variant .so> picks a
variant build.   python scripts/e2e_sweep.py [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pool = torch.randint(0, 4, (8, n), dtype=torch.uint8).pin_memory()
for fmt in ("bytes", "nibble"):
    for chunks in (1, 2, 3, 4, 6):
        h = g.HostSteppedEnv(n, seed=1, n_chunks=chunks, board_format=fmt)
        h.reset()
        for i in range(10):
            h.step_pinned(pool[i % 8])
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for i in range(100):
                h.step_pinned(pool[i % 8])
            best = min(best, (time.perf_counter() - t0) / 100)
        print("%s  format %-6s chunks %d  %.3f ms/step  %.3e steps/s" % (os.environ.get("G2048_SO", "default")[-24:], fmt, chunks, best * 1e3, n / best), flush=True)
        h.close()
