#!/usr/bin/env python
"""The step with every optional output (illegal, highest, legal mask, terminal boards, episode statistics)
against the lean step, device-resident, CUDA-event timed.  Run under gpurun."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402

dev = torch.device("cuda", 0)
for n in (1 << 20, 262144):
    for name, outputs in (("lean", ()), ("legal_mask", ("legal_mask",)), ("all", g.ALL_OUTPUTS)):
        games = [g.BatchedGame2048(n, seed=s, device=dev, env_id_base=s * n, outputs=outputs) for s in range(8)]
        acts = torch.randint(0, 4, (16, n), dtype=torch.uint8, device=dev)
        for gm in games:
            gm.reset()
            for t in range(6):
                gm.step(acts[t])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 3000
        e0.record()
        for t in range(K):
            games[t % 8].step(acts[t % 16])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / K
        print("n %8d outputs %-10s %.2f us/step  %.3e steps/s" % (n, name, us, n / us * 1e6))
        del games
