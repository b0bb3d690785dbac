#!/usr/bin/env python
"""g2048_env_step_host with the boards packed on the wire (G2048_BOARDS_BYTES_PACKED_WIRE) against the plain wire:
ms per step over slice counts and unpack-thread counts.   python scripts/e2e_wire_sweep.py [n] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402


def run(n, steps, **kw):
    h = g.HostSteppedEnv(n, seed=1, **kw)
    h.reset()
    pool = torch.randint(0, 4, (8, n), dtype=torch.uint8, generator=torch.Generator().manual_seed(7)).pin_memory()
    for i in range(10):
        h.step_pinned(pool[i % 8])
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(steps):
            h.step_pinned(pool[i % 8])
        dt = (time.perf_counter() - t0) / steps
        best = dt if best is None else min(best, dt)
    chk = int(h.buffers.boards.to(torch.int64).sum())
    h.close()
    return best * 1e3, chk


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    print("cpus of the process: %d" % len(os.sched_getaffinity(0)))
    ms, ref = run(n, steps, wire="plain", n_chunks=2)
    print("plain wire   chunks 2             %.3f ms/step  %.3e steps/s" % (ms, n / ms * 1e3), flush=True)
    for chunks in (0, 4, 5, 6):
        for threads in (0, 4, 12):
            ms, chk = run(n, steps, wire="packed", n_chunks=chunks, unpack_threads=threads)
            print("packed wire  chunks %-2d threads %-2d  %.3f ms/step  %.3e steps/s  %s" % (
                chunks, threads, ms, n / ms * 1e3, "boards identical" if chk == ref else "BOARDS DIFFER"), flush=True)


if __name__ == "__main__":
    main()
