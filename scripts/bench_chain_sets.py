#!/usr/bin/env python
"""Chained launches: direct vs interleaved launch shape as a function of how many env sets alternate on the stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gym_2048_b200 as g  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_chain import timed  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    steps = 3000
    for n in (1 << 20, 1 << 18, 1 << 17):
        gen = torch.Generator(device=dev).manual_seed(1)
        pool = torch.randint(0, 4, (16, n), generator=gen, device=dev, dtype=torch.uint8)
        for S in (1, 2, 3, 4, 6, 8, 16):
            row = []
            for mode in (False, True, "interleaved"):
                games = [g.BatchedGame2048(n, seed=42, device=dev, env_id_base=s * n, outputs=()) for s in range(S)]
                for gm in games:
                    gm.reset()
                row.append(min(timed(games, pool, steps, mode) for _ in range(2)))
                del games
            print("n %8d  %2d sets | plain %6.2f  chained direct %6.2f  chained interleaved %6.2f us per launch" % (n, S, *row), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    main()
