#!/usr/bin/env python
"""Chained launches (G2048_FLAG_CHAINED): bit-exactness against plain launches and launch time, per batch size.

    python scripts/bench_chain.py [sets] [steps]

For every size: S env sets stepped round-robin (the bench's pattern: consecutive launches are independent, each set
chains to its own launch S launches back) and ONE set stepped back to back (every launch waits for its predecessor's
slices), plain vs chained, through StepSchedule (one C call issues all launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gym_2048_b200 as g  # noqa: E402


def timed(games, pool, steps, chained, policy=None):
    sched = g.StepSchedule()
    S, P = len(games), pool.shape[0]
    for j in range(steps):
        if policy:
            sched.add(games[j % S], policy=policy, chained=chained)
        else:
            sched.add(games[j % S], pool[j % P], chained=chained)
    sched.build()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    warm = steps // 4
    sched.run(0, warm)
    e0.record()
    sched.run(warm, steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (steps - warm)


def main():
    sets = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    dev = torch.device("cuda", 0)
    sizes = [int(x) for x in os.environ.get("G2048_CHAIN_SIZES", "1048576,524288,262144,131072,65536").split(",")]
    for n in sizes:
        gen = torch.Generator(device=dev).manual_seed(1)
        pool = torch.randint(0, 4, (16, n), generator=gen, device=dev, dtype=torch.uint8)
        row = {}
        for S in (sets, 1):
            for outputs, policy, tag in (((), None, "lean"), (("legal_mask",), "legal", "c4")):
                final = {}
                for chained in (False, True):
                    games = [g.BatchedGame2048(n, seed=42, device=dev, env_id_base=s * n, outputs=outputs) for s in range(S)]
                    for gm in games:
                        gm.reset()
                    us = min(timed(games, pool, steps, chained and ("interleaved" if S > 1 else True), policy) for _ in range(2))
                    torch.cuda.synchronize()
                    final[chained] = (torch.stack([gm.boards for gm in games]).clone(), games[0].rewards.clone(),
                                      games[0]._dones.clone())
                    row[(S, tag, chained)] = us
                    del games
                same = all(torch.equal(a, b) for a, b in zip(final[False], final[True]))
                row[(S, tag, "same")] = same
        print("n %8d | %2d sets lean %6.2f -> %6.2f us %s | c4 %6.2f -> %6.2f us %s | 1 set lean %6.2f -> %6.2f us %s | c4 %6.2f -> %6.2f us %s" % (
            n, sets, row[(sets, "lean", False)], row[(sets, "lean", True)], "bit-exact" if row[(sets, "lean", "same")] else "MISMATCH",
            row[(sets, "c4", False)], row[(sets, "c4", True)], "bit-exact" if row[(sets, "c4", "same")] else "MISMATCH",
            row[(1, "lean", False)], row[(1, "lean", True)], "bit-exact" if row[(1, "lean", "same")] else "MISMATCH",
            row[(1, "c4", False)], row[(1, "c4", True)], "bit-exact" if row[(1, "c4", "same")] else "MISMATCH"), flush=True)


if __name__ == "__main__":
    main()
