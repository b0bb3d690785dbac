#!/usr/bin/env python
"""The step kernel's compile-time output sets against the lean step, device-resident, CUDA-event timed, issued
through g2048_step_list (StepSchedule) so that the host's launch rate does not bound the small sizes.
Run under gpurun:  python scripts/bench_extras.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402

dev = torch.device("cuda", 0)
BYTES = {"lean": 38, "legal_mask": 39, "evaluator": 41, "all": 38, "policy_legal": 40, "policy_uniform": 38}
for n in (1 << 20, 262144, 65536):
    S = 32 if n < (1 << 20) else 8
    for name, outputs, policy in (("lean", (), None), ("legal_mask", ("legal_mask",), None),
                                  ("evaluator", ("illegal", "highest", "legal_mask"), None), ("all", g.ALL_OUTPUTS, None),
                                  ("policy_uniform", (), "uniform"), ("policy_legal", ("legal_mask",), "legal")):
        games = [g.BatchedGame2048(n, seed=s, device=dev, env_id_base=s * n, outputs=outputs) for s in range(S)]
        acts = torch.randint(0, 4, (16, n), dtype=torch.uint8, device=dev)
        for gm in games:
            gm.reset()
            if "legal_mask" in outputs:
                gm.step_many(policy="legal", n_steps=150) if "episode" not in outputs else None     # mid-game boards
            for t in range(6):
                gm.step(acts[t])
        K = 3000
        sched = g.StepSchedule()
        outs = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(S)]
        for t in range(2 * K):
            if policy:
                sched.add(games[t % S], outs[t % S], policy=policy)
            else:
                sched.add(games[t % S], acts[t % 16])
        sched.run(0, K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sched.run(K, 2 * K)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / K
        print("n %8d  %-14s %6.2f us/step  %.3e steps/s  %4.2f of the HBM roofline at %d B/board" % (
            n, name, us, n / us * 1e6, BYTES[name] * n / us / 1e3 / 6455.9, BYTES[name]), flush=True)
        del games, sched
        torch.cuda.empty_cache()
