#!/usr/bin/env python
"""Host-side cost of one BatchedGame2048.step() call (tiny batch, so the GPU never limits):
wall time per call for the public method and for its pieces.  Run under gpurun."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402


def per_call(fn, iters=50000):
    for _ in range(2000):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    dt = (time.perf_counter() - t0) / iters
    torch.cuda.synchronize()
    return dt * 1e6


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    for outputs in ((), g.ALL_OUTPUTS):
        game = g.BatchedGame2048(1024, seed=1, device=dev, outputs=outputs)
        game.reset()
        act = torch.zeros(1024, dtype=torch.uint8, device=dev)
        print("outputs=%s: step() %.2f us/call" % ("lean" if not outputs else "all", per_call(lambda: game.step(act))))
    lib = game.lib
    args, ref = game._args, game._args_ref
    stream = torch.cuda.current_stream(dev).cuda_stream
    print("raw ctypes g2048_step            %.2f us/call" % per_call(lambda: lib.g2048_step(ref, stream)))
    print("torch.cuda.current_stream().cuda_stream %.2f us" % per_call(lambda: torch.cuda.current_stream(dev).cuda_stream))
    print("torch._C._cuda_getCurrentRawStream %.2f us" % per_call(lambda: torch._C._cuda_getCurrentRawStream(0)))
    print("torch.cuda.current_device()      %.2f us" % per_call(torch.cuda.current_device))
    print("act.data_ptr()                   %.2f us" % per_call(act.data_ptr))
    print("sample_actions()                 %.2f us" % per_call(lambda: game.sample_actions(out=act)))
    print("observe(u8)                      %.2f us" % per_call(lambda: game.observe(torch.uint8), 20000))


if __name__ == "__main__":
    main()
