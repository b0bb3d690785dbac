#!/usr/bin/env python
"""Launch one step-kernel specialisation repeatedly, for ncu:  python scripts/profile_kernels.py <lean|chain|chain_direct|mask|all|eprun|eval> [n] [launches]
(8 env sets stepped round-robin so the working set exceeds the L2 at 1 Mi boards.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402

SETS = {"lean": (), "chain": (), "chain_direct": (), "mask": ("legal_mask",), "all": g.ALL_OUTPUTS, "eprun": None, "eval": ("illegal", "highest", "legal_mask")}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "lean"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    launches = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    pool = torch.randint(0, 4, (8, n), generator=gen, device=dev, dtype=torch.uint8)
    if which == "eprun":                                  # the kernel behind g2048_env_step_host (bench.py's e2e leg)
        h = g.HostSteppedEnv(n, seed=1, n_chunks=1)
        h.reset()
        host = pool.cpu().pin_memory()
        for j in range(launches):
            h.step_pinned(host[j % 8])
        return
    games = [g.BatchedGame2048(n, seed=1, device=dev, env_id_base=s * n, outputs=SETS[which]) for s in range(8)]
    for gm in games:
        gm.reset()
    if which == "mask":                                   # config 4: mid-game boards under the random-legal policy
        for gm in games:
            gm.step_many(policy="legal", n_steps=200)
    if which.startswith("chain"):                         # the benchmark's launches: chained, issued by one C call
        if which == "chain_direct":
            games = games[:1]
        sched = g.StepSchedule()
        for j in range(launches):
            sched.add(games[j % len(games)], pool[j % 8], chained="interleaved" if which == "chain" else True)
        sched.run()
        torch.cuda.synchronize()
        return
    for j in range(launches):
        games[j % 8].step(pool[j % 8])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
