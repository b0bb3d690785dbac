#!/usr/bin/env python
"""Benchmark of the batched 2048 step path (BASELINE.json metric: env steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours (default): one process per GPU (torchrun for N>1, NCCL only for the barrier and the
max-over-ranks reduction — the step path has no collective).  A "step" is one call of
BatchedGame2048.step() over one batch of `--envs` boards per GPU (default 1,048,576 =
BASELINE config 3) with uniform-random actions; `--sets` independent batches are stepped
round-robin so the working set exceeds the 126 MB L2.  Prints ONE JSON line with `value`
(device-resident throughput), `e2e` (host-buffer C-ABI call, copies inside the timed
region), `roofline`, `cpu_baseline`, `clocks`, `gpu_launches`, and `fused`: the same workload through
g2048_step_many (`--fused-steps` steps per launch, boards in registers in between) — an extra, not the
headline.

reference: the UNMODIFIED reference env's step() (baseline/_ref, installed by
__graft_entry__.build()) on all host cores, one env per process, same action distribution;
falls back to the multi-threaded C oracle port when the reference package is absent.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
ALG_BYTES_PER_STEP = 38          # board in 16 + action 1 + board out 16 + reward 4 + done 1
FALLBACK_HBM_GBS = 6650.0        # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
# stdout at the first collective whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for
# the whole run and the JSON line goes to the saved original.
_REAL_STDOUT = None


def guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


# ----------------------------------------------------------------------------- reference arm
def _ref_paths():
    paths = []
    try:
        import gymnasium  # noqa: F401
    except ImportError:
        paths.append(os.path.join(ROOT, "oracle", "shim"))
    paths.append(os.path.join(ROOT, "baseline", "_ref"))
    return paths


def reference_available():
    return os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "env", "envs", "game2048_env.py"))


def _ref_worker(args):
    """One process, one reference env: `n_steps` step() calls with uniform-random actions,
    reset() on terminated.  Returns (steps, seconds)."""
    idx, n_steps, warm = args
    for p in _ref_paths():
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    from env.envs.game2048_env import Game2048Env       # the reference, unmodified
    env = Game2048Env()
    env.reset(seed=idx)
    acts = np.random.default_rng(1000 + idx).integers(0, 4, size=n_steps + warm)
    t0 = 0.0
    for i in range(n_steps + warm):
        if i == warm:
            t0 = time.perf_counter()
        _, _, terminated, _, _ = env.step(int(acts[i]))
        if terminated:
            env.reset()
    return n_steps, time.perf_counter() - t0


def time_reference(procs, steps_per_proc, warm=200, pool=None):
    """Aggregate steps/s of `procs` reference envs stepping concurrently."""
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(procs)
    try:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(i, steps_per_proc, warm) for i in range(procs)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest, wall


def time_port(n_envs, steps, threads, seed=42):
    """The C oracle port (oracle/g2048_oracle.c), multi-threaded, same workload shape."""
    import numpy as np
    from oracle import oracle
    b = oracle.OracleBatch(n_envs, seed=seed, threads=threads)
    b.reset()
    import ctypes as C
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 4, (steps + 1, n_envs)).astype(np.uint8)
    rewards, dones = np.zeros(n_envs, np.float32), np.zeros(n_envs, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)            # noqa: E731

    def one(t):
        a = oracle.StepArgs()
        a.boards, a.actions, a.rewards, a.dones = p(b.boards), p(acts[t]), p(rewards), p(dones)
        a.n, a.seed, a.step_index, a.flags = n_envs, seed, t, 1
        assert oracle.lib().g2048_oracle_step_mt(C.byref(a), threads) == 0
    one(0)
    t0 = time.perf_counter()
    for t in range(1, steps + 1):
        one(t)
    return n_envs * steps / (time.perf_counter() - t0)


def host_cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    """`--impl reference`: the UNMODIFIED reference step() on all host cores.  A bench "step" is a
    bounded sample of the workload: `per` step() calls on each of `cores` single-env processes.  All
    W + K bench steps of a worker run inside ONE process-pool call (no per-step IPC), and `per` shrinks
    as K grows, so the whole run takes seconds to a few minutes whatever K the driver passes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    config = workload_config(args, world=args.gpus)
    K, W = max(args.steps, 1), max(args.warmup, 0)
    if reference_available():
        kind = "reference"
        budget_calls = 400000                                    # per process: ~20 s at ~20k step()/s
        per = max(1, min(args.ref_steps_per_proc, budget_calls // (K + W)))
        value, wall = time_reference(cores, per * K, warm=per * W + 50)
        dt = cores * per * K / value
        sample = "%d processes x 1 unmodified reference env x %d step() calls per bench step, uniform-random " \
                 "actions, reset() on terminated" % (cores, per)
    else:
        kind = "port"
        n = 1 << 18
        for _ in range(min(W, 3)):
            time_port(n, 2, cores)
        t0 = time.perf_counter()
        reps = min(K, 64)
        for _ in range(reps):
            time_port(n, 4, cores)
        dt = (time.perf_counter() - t0) * K / reps
        value = n * 4 * K / dt
        sample = "C oracle port, %d threads, %d envs x 4 steps per bench step" % (cores, n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "cpu": host_cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- our arm
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled in a separate process during the timed region."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
             "clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_start, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        inside = [r for t, r in self.rows if t_start <= t <= t_end] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, world):
    per_gpu = args.envs if args.scaling == "weak" else args.envs // world
    return {"workload": "BASELINE config 3: %d parallel 4x4 envs per GPU, uniform-random actions, auto-reset"
                        % per_gpu,
            "envs_per_gpu": per_gpu, "global_envs": per_gpu * world, "parallelism": "independent slices x%d" % world,
            "l2_policy": "%d independent env sets stepped round-robin (working set > 126 MB L2)" % args.sets,
            "sets": args.sets, "outputs": "boards,rewards,dones (lean kernel)"}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def measured_traffic(n):
    """DRAM bytes per launch of the step kernel from the committed ncu capture, scaled to n."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            t = json.load(f)
        return t["dram_bytes_per_board"] * n
    except (OSError, KeyError, ValueError):
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gym_2048_b200 as g

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.envs if args.scaling == "weak" else args.envs // world
    total_envs = n * world
    K, W, R = args.steps, args.warmup, args.sets
    # R independent batches; global env ids are disjoint across sets and ranks
    games = [g.BatchedGame2048(n, seed=42, device=dev, env_id_base=s * total_envs + rank * n, outputs=())
             for s in range(R)]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = torch.randint(0, 4, (args.action_pool, n), generator=gen, device=dev, dtype=torch.uint8)
    for gm in games:                       # boards generated by the env itself: reset + a few real steps
        gm.reset()
        for t in range(6):
            gm.step(pool[t % args.action_pool])
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(W):
        games[i % R].step(pool[i % args.action_pool])
    torch.cuda.synchronize()
    if sampler:
        sampler.wait_first_sample()
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    start.record()
    for i in range(K):
        games[i % R].step(pool[i % args.action_pool])
    end.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    barrier()
    dev_ms = start.elapsed_time(end)
    ms = max_over_ranks(dev_ms)
    clocks = sampler.stop(t_start, t_end) if sampler else None
    value = total_envs * K / (ms * 1e-3)
    launch_s = ms * 1e-3 / K              # the region holds only step-kernel launches, back to back

    # ---- several steps per launch (g2048_step_many): the same steps with the boards held in registers in
    #      between — the open-loop form of this workload (the actions are pre-generated).  Reported beside the
    #      headline, never instead of it: a step here moves 6 bytes, not 38.
    fused = None
    if args.fused_steps > 0:
        Kf = args.fused_steps
        facts = torch.randint(0, 4, (Kf, n), generator=gen, device=dev, dtype=torch.uint8)
        frew = torch.empty((Kf, n), dtype=torch.float32, device=dev)
        fdone = torch.empty((Kf, n), dtype=torch.uint8, device=dev)
        launches = max(2, min(200, (1 << 31) // (Kf * n)))
        for i in range(2):
            games[i % R].step_many(facts, rewards=frew, dones=fdone)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record()
        for i in range(launches):
            games[i % R].step_many(facts, rewards=frew, dones=fdone)
        f1.record()
        torch.cuda.synchronize()
        f_ms = max_over_ranks(f0.elapsed_time(f1))
        fused = {"api": "g2048_step_many", "steps_per_launch": Kf, "launches": launches,
                 "value": total_envs * Kf * launches / (f_ms * 1e-3), "unit": UNIT,
                 "us_per_step": f_ms * 1e3 / (Kf * launches),
                 "algorithmic_bytes_per_step": 6 + 32.0 / Kf,
                 "note": "bit-identical to steps_per_launch calls of step(); boards stay in registers between steps"}
        del facts, frew, fdone

    # ---- e2e: host buffers through the C-ABI handle, copies inside the timed region ----------
    Ke = max(3, min(K, args.e2e_steps))
    henv = g.HostSteppedEnv(n, seed=42, device=local, env_id_base=rank * n, n_chunks=args.e2e_chunks)
    henv.reset()
    host_pool = torch.randint(0, 4, (8, n), dtype=torch.uint8).pin_memory()
    for i in range(3):
        henv.step_pinned(host_pool[i % 8])
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        henv.step_pinned(host_pool[i % 8])
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    chk = float(henv.buffers.rewards.sum())        # the result is read on the host
    e2e_value = total_envs * Ke / e2e_s
    e2e_launches = henv.n_chunks_effective * Ke
    henv.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = hbm_peak()
    achieved = ALG_BYTES_PER_STEP * n / launch_s / 1e9
    traffic = measured_traffic(n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "g2048_step_kernel<false, false>",
                     "algorithmic_bytes_per_launch": ALG_BYTES_PER_STEP * n,
                     "launch_us": launch_s * 1e6, "frac_of_nominal_8TBs": achieved / 8000.0},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": n * 21,
                "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke, "chunks": args.e2e_chunks,
                "api": "g2048_env_step_host (pinned host buffers)", "checksum": chk},
        "fused": fused,
        "gpu_launches": K, "e2e_gpu_launches": e2e_launches,
        "clocks": clocks,
        "gpu": torch.cuda.get_device_name(local),
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cb = {}
        if reference_available():
            per = args.ref_steps_per_proc * 30          # ~15 s of CPU work per core's worth of processes
            v, wall = time_reference(cores, per)
            cb = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                  "sample": "%d processes x 1 unmodified reference env x %d step() calls, uniform-random "
                            "actions, reset() on terminated" % (cores, per), "wall_s": wall}
        port = time_port(1 << 18, 8, cores)
        if not cb:
            cb = {"value": port, "unit": UNIT, "cores": cores, "kind": "port",
                  "sample": "C oracle port, %d threads, 262144 envs x 8 steps" % cores}
        cb["cpu"] = host_cpu_model()
        cb["c_port_value"] = port
        line["cpu_baseline"] = cb
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=1 << 20, help="boards per GPU (weak) or in total (strong)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--sets", type=int, default=8)
    ap.add_argument("--action-pool", type=int, default=16)
    ap.add_argument("--fused-steps", type=int, default=32, help="steps per g2048_step_many launch (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--e2e-chunks", type=int, default=3)
    ap.add_argument("--ref-steps-per-proc", type=int, default=5000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        args.steps = 20 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 50000 if args.steps is None else args.steps
    args.warmup = 200 if args.warmup is None else max(args.warmup, 3)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
