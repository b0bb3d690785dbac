#!/usr/bin/env python
"""Small shards: who bounds a step, the host's launch rate or the GPU?  For n boards per launch prints
  host  us per launch the C loop (g2048_step_list) needs to ISSUE K launches (perf_counter around the call, no sync)
  gpu   us per launch the GPU needs (CUDA events, back to back), for: step_list over 32 sets (HBM-resident),
        step_n on one set (L2-resident), a CUDA graph of 64 captured steps (device-side step index), Python step()
Run on the GPU box: python scripts/launch_rate.py [n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [32768, 65536, 131072, 262144, 1 << 20]
    dev = torch.device("cuda", 0)
    for n in sizes:
        S, P, K = 32, 16, 4000
        games = [g.BatchedGame2048(n, seed=1, device=dev, env_id_base=s * n, outputs=()) for s in range(S)]
        gen = torch.Generator(device=dev).manual_seed(0)
        pool = torch.randint(0, 4, (P, n), generator=gen, device=dev, dtype=torch.uint8)
        for gm in games:
            gm.reset()
        out = {}
        # step_list over S sets
        sched = g.StepSchedule()
        for j in range(2 * K):
            sched.add(games[j % S], pool[j % P])
        sched.build()
        sched.run(0, K)                      # warm
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        sched.run(K, 2 * K)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        out["list host"] = (t1 - t0) / K * 1e6
        out["list gpu"] = e0.elapsed_time(e1) / K * 1e3
        # Python loop
        for j in range(200):
            games[j % S].step(pool[j % P])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for j in range(K):
            games[j % S].step(pool[j % P])
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        out["py host"] = (t1 - t0) / K * 1e6
        out["py gpu"] = e0.elapsed_time(e1) / K * 1e3
        # step_n on one set (K launches, same boards: L2-resident for small n)
        acts = pool.repeat(K // P, 1).contiguous()
        rew = torch.empty((K, n), dtype=torch.float32, device=dev) if K * n * 5 < (8 << 30) else None
        if rew is not None:
            don = torch.empty((K, n), dtype=torch.uint8, device=dev)
            games[0].step_n(acts[:64], rewards=rew[:64], dones=don[:64])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            games[0].step_n(acts, rewards=rew, dones=don)
            t1 = time.perf_counter()
            e1.record()
            torch.cuda.synchronize()
            out["step_n host"] = (t1 - t0) / K * 1e6
            out["step_n gpu"] = e0.elapsed_time(e1) / K * 1e3
            del rew, don
        # CUDA graph of 64 steps round-robin over the sets (device-side step index per game)
        T = 64

        def body():
            for j in range(T):
                games[j % S].step(pool[j % P])
        for gm in games:
            gm.use_device_step_counter(True)
        replay = games[0].capture(body, warmup=1)
        for _ in range(5):
            replay()
        torch.cuda.synchronize()
        reps = max(K // T, 10)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            replay()
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        out["graph host"] = (t1 - t0) / (reps * T) * 1e6
        out["graph gpu"] = e0.elapsed_time(e1) / (reps * T) * 1e3
        print("n %8d  " % n + "  ".join("%s %.2f" % kv for kv in out.items()), flush=True)
        del games, pool, acts, sched
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
