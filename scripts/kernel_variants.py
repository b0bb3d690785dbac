#!/usr/bin/env python
"""Build (here, nvcc) and time (on the GPU box) compile-time variants of the step kernel.

  python scripts/kernel_variants.py build  name:-DFLAG=V,-DFLAG2=V ...   # in the container
  python scripts/kernel_variants.py run [n_envs] [steps]                  # under gpurun

Each variant is a separate libg2048 build under gym-2048_b200/variants/ (git-ignored .so);
`run` checks every variant against variant `base` bit for bit before timing it."""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "gym-2048_b200", "variants")


def build(specs):
    from gym_2048_b200 import _lib
    os.makedirs(VDIR, exist_ok=True)
    for f in glob.glob(os.path.join(VDIR, "*.so")):
        os.remove(f)
    for spec in specs:
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        out = os.path.join(VDIR, "libg2048_%s.so" % name)
        cmd = ["nvcc"] + _lib.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", out] + _lib.SOURCES
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(r.stderr)
            raise SystemExit(1)
        lines = r.stderr.splitlines()
        info = [lines[i + 2].strip() for i, l in enumerate(lines) if "step_kernelILj0ELb0ELi0E" in l and "Compiling" in l]
        print(name, flags, info)


def run(n=1 << 20, steps=3000, sets=8):
    sets = int(os.environ.get("G2048_VARIANT_SETS", sets))
    import torch
    from gym_2048_b200._lib import StepArgs
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gen = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, 4, (16, n), generator=gen, device=dev, dtype=torch.uint8)
    results = {}
    ref_boards = None
    names = sorted(os.path.basename(f)[9:-3] for f in glob.glob(os.path.join(VDIR, "*.so")))
    if "base" in names:
        names.remove("base")
        names.insert(0, "base")
    for name in names:
        L = C.CDLL(os.path.join(VDIR, "libg2048_%s.so" % name))
        L.g2048_step.argtypes = [C.POINTER(StepArgs), C.c_void_p]
        L.g2048_reset.argtypes = [C.c_void_p] * 2 + [C.c_uint64] * 4 + [C.c_void_p]
        L.g2048_last_error.restype = C.c_char_p
        boards = [torch.zeros((n, 16), dtype=torch.uint8, device=dev) for _ in range(sets)]
        rewards = torch.zeros(n, dtype=torch.float32, device=dev)
        dones = torch.zeros(n, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        for s, b in enumerate(boards):
            assert L.g2048_reset(b.data_ptr(), None, n, s * n, 42, 0, stream) == 0
        args = StepArgs()
        args.rewards, args.dones, args.n, args.seed, args.flags = rewards.data_ptr(), dones.data_ptr(), n, 42, 1

        def step(t):
            args.boards = boards[t % sets].data_ptr()
            args.actions = acts[t % 16].data_ptr()
            args.env_id_base = (t % sets) * n
            args.step_index = t
            rc = L.g2048_step(C.byref(args), stream)
            assert rc == 0, L.g2048_last_error()
        for t in range(64):
            step(t)
        torch.cuda.synchronize()
        snap = torch.stack([b.clone() for b in boards])
        if ref_boards is None:
            ref_boards = snap
        ok = bool(torch.equal(snap, ref_boards))
        best = None
        # small batches: a Python loop cannot issue launches fast enough — pre-build the calls, one C call issues them
        use_list = n < (1 << 19) and hasattr(L, "g2048_step_list")
        if use_list:
            L.g2048_step_list.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        for rep in range(3):
            if use_list:
                arr = (StepArgs * steps)()
                for j, t in enumerate(range(64 + rep * steps, 64 + (rep + 1) * steps)):
                    C.memmove(C.byref(arr[j]), C.byref(args), C.sizeof(StepArgs))
                    arr[j].boards = boards[t % sets].data_ptr()
                    arr[j].actions = acts[t % 16].data_ptr()
                    arr[j].env_id_base = (t % sets) * n
                    arr[j].step_index = t
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if use_list:
                assert L.g2048_step_list(C.byref(arr), steps, stream) == 0, L.g2048_last_error()
            else:
                for t in range(64 + rep * steps, 64 + (rep + 1) * steps):
                    step(t)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / steps
            best = us if best is None else min(best, us)
        results[name] = best
        print("%-24s %8.2f us/step  %.3e steps/s  %s" % (name, best, n / best * 1e6, "bit-exact" if ok else "MISMATCH"),
              flush=True)
    return results


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(*[int(x) for x in sys.argv[2:]])
