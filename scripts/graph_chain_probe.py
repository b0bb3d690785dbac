#!/usr/bin/env python
"""Probe: do CHAINED launches keep their overlap when they are replayed from a CUDA graph (no host cost per launch)?
K chained launches over S env sets, captured once (step indices baked in: the graph is valid for ONE replay), replayed,
timed with two events; against the same launches issued by g2048_step_list from one host thread."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gym_2048_b200 as g  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    S, K = 32, 2048
    for n in (131072, 262144, 1 << 20):
        gen = torch.Generator(device=dev).manual_seed(1)
        pool = torch.randint(0, 4, (16, n), generator=gen, device=dev, dtype=torch.uint8)
        res = {}
        for mode in ("list", "graph"):
            for chained in (False, "interleaved"):
                games = [g.BatchedGame2048(n, seed=42, device=dev, env_id_base=s * n, outputs=()) for s in range(S)]
                for gm in games:
                    gm.reset()
                warm = g.StepSchedule()
                for j in range(256):
                    warm.add(games[j % S], pool[j % 16], chained=chained)
                warm.run()
                sched = g.StepSchedule()
                for j in range(K):
                    sched.add(games[j % S], pool[j % 16], chained=chained)
                sched.build()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if mode == "list":
                    e0.record()
                    sched.run()
                    e1.record()
                else:
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.stream(side):
                        with torch.cuda.graph(graph, stream=side):
                            sched.run()
                    torch.cuda.current_stream(dev).wait_stream(side)
                    torch.cuda.synchronize()
                    e0.record()
                    graph.replay()
                    e1.record()
                torch.cuda.synchronize()
                res[(mode, bool(chained))] = (e0.elapsed_time(e1) * 1e3 / K, torch.stack([gm.boards for gm in games]).clone())
                del games
        same = torch.equal(res[("list", False)][1], res[("graph", True)][1]) and torch.equal(res[("list", False)][1], res[("list", True)][1])
        print("n %8d | list: plain %.2f chained %.2f us | graph replay: plain %.2f chained %.2f us per launch | %s" % (
            n, res[("list", False)][0], res[("list", True)][0], res[("graph", False)][0], res[("graph", True)][0],
            "bit-exact" if same else "MISMATCH"), flush=True)


if __name__ == "__main__":
    main()
