#!/usr/bin/env python
"""A RANGE of overlapping chained launches under `ncu --replay-mode range` (cudaProfilerStart/Stop around one
g2048_step_list call): pipe utilisation of the steady state, which a per-kernel capture cannot see (ncu serialises
kernels; chained launches only reach their speed when several share the machine).
    ncu --replay-mode range --metrics ... python scripts/profile_range.py [chained|plain|c4|c4_plain] [n] [launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "chained"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    launches = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    pool = torch.randint(0, 4, (8, n), generator=gen, device=dev, dtype=torch.uint8)
    S = 32
    c4 = mode.startswith("c4")                   # BASELINE config 4: legal mask + random-legal policy drawn in the kernel
    games = [g.BatchedGame2048(n, seed=1, device=dev, env_id_base=s * n, outputs=("legal_mask",) if c4 else ()) for s in range(S)]
    for gm in games:
        gm.reset()
        if c4:
            gm.step_many(policy="legal", n_steps=150)        # mid-game boards
    chained = False if mode.endswith("plain") else "interleaved"
    warm, sched = g.StepSchedule(), g.StepSchedule()
    for j in range(128):
        warm.add(games[j % S], None if c4 else pool[j % 8], policy="legal" if c4 else None, chained=chained)
    for j in range(launches):
        sched.add(games[j % S], None if c4 else pool[j % 8], policy="legal" if c4 else None, chained=chained)
    sched.build()
    warm.run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    sched.run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
