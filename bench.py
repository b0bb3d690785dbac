#!/usr/bin/env python
"""Benchmark of the batched 2048 step path (BASELINE.json metric: env steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours (default): one process per GPU (torchrun for N>1, NCCL only for the barrier and the
max-over-ranks reduction — the step path has no collective).  A "step" is ONE KERNEL LAUNCH of
g2048_step over one batch of boards — `--envs` (default 1,048,576 = BASELINE config 3) in total,
sharded over the N ranks — with uniform-random actions; `--sets` independent batches are stepped
round-robin so the working set exceeds the 126 MB L2.  The launches are CHAINED launches
(include/g2048.h, G2048_FLAG_CHAINED | _CHAIN_INTERLEAVED: each depends on the previous step of its
own env set warp by warp instead of on the whole previous grid; `--chain off` for plain ones, which
the `plain_launches` block times anyway) issued through g2048_step_list by `--issue-threads` host
threads (auto: two, each with its own stream and env sets).  Prints ONE JSON line with `value`
(device-resident throughput), `e2e` (host-buffer C-ABI call, copies inside the timed
region), `roofline`, `cpu_baseline`, `clocks`, `gpu_launches`, `long_region` (regions of
`--long-region` launches: an event every 1000 launches instead of every K), and `fused`: the same workload through
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
    K, W = max(args.steps, 1), max(args.warmup, 0)
    if reference_available():
        kind = "reference"
        budget_calls = 400000                                    # per process: ~20 s at ~20k step()/s
        per = max(1, min(args.ref_steps_per_proc, budget_calls // (K + W)))
        value, wall = time_reference(cores, per * K, warm=per * W + 50)
        dt = cores * per * K / value
        sample = "%d processes x 1 unmodified reference env x %d step() calls per bench step, uniform-random " \
                 "actions, reset() on terminated" % (cores, per)
        ran = {"workload": "%d single 4x4 reference envs (unmodified Game2048Env.step incl. stack()), one per host "
                           "core, uniform-random actions, reset() on terminated — the reference has no batched form" % cores,
               "envs": cores, "envs_per_process": 1, "processes": cores, "step_calls_per_bench_step": cores * per,
               "parallelism": "multiprocessing, one env per process"}
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
        ran = {"workload": "C oracle port of the reference step(), %d envs stepped by %d threads" % (n, cores),
               "envs": n, "threads": cores}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        # `config` names the benchmark both arms are quoted on (the GPU arm's: BASELINE config 3); what THIS arm
        # executed on the host cores — it has no batched form — is `reference_workload`
        "config": workload_config(args, world=args.gpus), "reference_workload": ran,
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


def workload_config(args, world, scaling=None, issue=None):
    scaling = scaling or args.scaling
    per_gpu = args.envs if scaling == "weak" else args.envs // world
    cfg = {"workload": "BASELINE config 3: %d parallel 4x4 envs in total, sharded as %d contiguous slices of %d "
                       "(one per GPU), uniform-random actions, auto-reset" % (per_gpu * world, world, per_gpu),
           "envs_per_gpu": per_gpu, "global_envs": per_gpu * world, "parallelism": "independent slices x%d" % world,
           "l2_policy": "%d independent env sets stepped round-robin, one launch per set per step: %.0f MB of "
                        "algorithmic traffic per GPU between two visits of a set (> 126 MB L2)"
                        % (args.sets, args.sets * per_gpu * ALG_BYTES_PER_STEP / 1e6),
           "sets": args.sets, "outputs": "boards,rewards,dones (lean kernel)"}
    if issue:
        cfg["issue"] = issue
    return cfg


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def committed_ncu_traffic(n):
    """DRAM bytes per launch of the step kernel from the COMMITTED ncu capture (profiles/step_kernel_traffic.json),
    scaled to n boards — a static figure, not a measurement of this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            t = json.load(f)
        return t["dram_bytes_per_board"] * n, t.get("source", "profiles/step_kernel_traffic.json"), \
            t.get("kind", "dram__bytes_read.sum + dram__bytes_write.sum, ncu (cold cache, serialised launches)")
    except (OSError, KeyError, ValueError):
        return None, None, None


def state_checksum(torch, dist, games, total_envs, rank, n, world, dev):
    """64-bit checksum of every board of every env set, keyed by the GLOBAL env id: independent of how the batch
    is sharded (sum over ranks), so a strong-scaling run must print the same value for N = 1, 2, 4, 8."""
    acc = torch.zeros(2, dtype=torch.int64, device=dev)
    ids = torch.arange(n, dtype=torch.int64, device=dev)
    for s, gm in enumerate(games):
        w = gm.boards.view(torch.int64).view(n, 2)
        gid = ids + (s * total_envs + rank * n)
        h = ((w[:, 0] * -7046029254386353131) ^ w[:, 1]) * (2 * gid + 1)          # int64 arithmetic wraps
        acc[0] += (h & 0xFFFFFFFF).sum()
        acc[1] += ((h >> 32) & 0xFFFFFFFF).sum()
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    lo, hi = int(acc[0].item()), int(acc[1].item())
    return "%016x" % ((lo + (hi << 32)) & (2**64 - 1))


def _add(sched, game, row, chained, policy):
    """One launch of a schedule: game.step(row), or — policy given — game.step(policy=...) writing its actions to row."""
    if policy:
        sched.add(game, row, policy=policy, chained=chained)
    else:
        sched.add(game, row, chained=chained)


def time_step_regions(torch, g, games, pool, spinup, W, K, R, small, barrier, max_over_ranks_vec, chained="interleaved",
                      policy=None):
    """Spin-up, warm-up and R back-to-back timed regions of K step launches each, round-robin over `games`.
    chained: how every launch depends on the previous step of ITS env set (BatchedGame2048.step): "interleaved" =
    chained launches (the action rows exist before the first launch and nothing else touches the sets, which is the
    promise the flag makes), False = plain launches.
    Returns (region times in ms, max over ranks, per region; issue description; t_start, t_end)."""
    S, P = len(games), pool.shape[0]
    total = spinup + W + R * K
    how = "chained launches (G2048_FLAG_CHAINED | G2048_FLAG_CHAIN_INTERLEAVED)" if chained else "plain launches"
    if small:
        # a shard the GPU steps faster than Python can launch: the whole schedule is built up front and each
        # region is ONE C call that issues its K launches (g2048_step_list)
        sched = g.StepSchedule()
        for j in range(total):
            _add(sched, games[j % S], pool[j % P], chained, policy)
        sched.build()
        issue = "g2048_step_list_timed via StepSchedule (all regions in one C call; one kernel launch per env step); " + how

        def run(lo, hi):
            sched.run(lo, hi)
    else:
        issue = "BatchedGame2048.step() from a Python loop (one kernel launch per env step); " + how

        def run(lo, hi):
            for j in range(lo, hi):
                games[j % S].step(pool[j % P], chained=chained)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    for e in ev:
        e.record()                                 # creates the cudaEvent_t handles (torch does it lazily)
    torch.cuda.synchronize()
    barrier()
    # no host sleep, synchronisation or barrier between the spin-up, the warm-up and the timed regions: the
    # clocks and the launch pipeline are in steady state when the first event is recorded
    run(0, spinup)
    run(spinup, spinup + W)
    t_start = time.perf_counter()
    at = spinup + W
    if small:
        # all R regions from ONE C call, the events recorded between them by the library: ~10 us of interpreter per
        # region would otherwise be a fifth of a region of twenty 3-us launches
        sched.run(at, at + R * K, events=ev, every=K)
    else:
        ev[0].record()
        for r in range(R):
            run(at, at + K)
            at += K
            ev[r + 1].record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    barrier()
    ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(R)]
    return max_over_ranks_vec(ms), issue, t_start, t_end


SPAN = {"ms_per_step": None}      # time_step_regions_mt's cross-check of its last call (see there)


def issue_plan(total, S, P, T, spinup_w, R, Kt):
    """Who launches what when T host threads share the single-stream schedule `launch j steps env set j % S with action
    row j % P, j < total`: thread t owns the sets t, t+T, ... and steps them round-robin, every set exactly as often
    and with exactly the rows the single stream would (so the boards do not depend on T).  Returns per thread
    (items, start): items = [(set, row), ...] in issue order, start = index of the first timed launch — R regions of
    Kt launches follow it; the starts are staggered by Kt/T launches from thread to thread (taken out of the spin-up),
    and at least max(64, 4 Kt) launches stay behind the last region where the schedule is long enough."""
    plan = []
    for t in range(T):
        mine = list(range(t, S, T))
        count = {s_: (total - s_ + S - 1) // S if s_ < total else 0 for s_ in mine}   # launches of set s_ in the single-stream order
        items, m = [], 0
        while any(m < count[s_] for s_ in mine):
            items += [(s_, (s_ + m * S) % P) for s_ in mine if m < count[s_]]
            m += 1
        tail = max(64, 4 * Kt)
        start = max(0, min(spinup_w // T, len(items) - R * Kt - tail) - ((T - 1 - t) * Kt) // T)
        plan.append((items, start))
    return plan


def time_step_regions_mt(torch, g, games, pool, spinup, W, K, R, T, barrier, max_over_ranks_vec, chained="interleaved",
                         policy=None):
    """The same launches as time_step_regions — global launch j steps env set j % S with action row j % P — issued by T
    host threads on T streams: thread t owns the sets t, t+T, ... (a set's launches stay on one stream, in order),
    and a timed region is K/T launches on each stream.  One host thread issues a launch every ~2.4 us on this box,
    two ~1.4 us between them (scripts/micro/launch_cost.cu): a 131,072-board shard, which the GPU steps in ~1.3 us,
    is otherwise timed at the speed of cudaLaunchKernelEx.  Region r's time is the slowest stream's (then the max
    over ranks); the streams' events are staggered by half a region so that they do not drain together.
    Returns (region times in ms; issue description; t_start, t_end; launches inside the timed regions)."""
    S, P = len(games), pool.shape[0]
    assert S % T == 0 and K % T == 0, "--sets and --steps must be multiples of the issuing threads"
    total, Kt = spinup + W + R * K, K // T
    dev = games[0].device
    streams = [torch.cuda.Stream(device=dev) for _ in range(T)]
    scheds, starts, evs = [], [], []
    for t, (items, start) in enumerate(issue_plan(total, S, P, T, spinup + W, R, Kt)):
        sched = g.StepSchedule()
        for s_, row in items:
            _add(sched, games[s_], pool[row], chained, policy)
        sched.build()
        assert start + R * Kt <= len(sched)
        scheds.append(sched)
        starts.append(start)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
        with torch.cuda.stream(streams[t]):
            for e in ev:
                e.record()
        evs.append(ev)
    torch.cuda.synchronize()
    barrier()
    cur = torch.cuda.current_stream(dev)
    errors = []
    go = threading.Barrier(T)

    def issue(t):
        try:
            with torch.cuda.device(dev), torch.cuda.stream(streams[t]):
                streams[t].wait_stream(cur)
                go.wait()                                             # the threads start issuing together
                scheds[t].run(0, starts[t])
                scheds[t].run(starts[t], starts[t] + R * Kt, events=evs[t], every=Kt)
                scheds[t].run()
        except Exception as e:                                        # noqa: BLE001 — re-raised on the main thread
            errors.append(e)
            go.abort()                                                # the other threads must not wait for this one
    t_start = time.perf_counter()
    threads = [threading.Thread(target=issue, args=(t,)) for t in range(T)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for st in streams:
        cur.wait_stream(st)
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    if errors:
        raise errors[0]
    barrier()
    ms = [max(evs[t][r].elapsed_time(evs[t][r + 1]) for t in range(T)) for r in range(R)]
    # cross-check: all timed launches over the device time from the first stream's first event to the last stream's
    # last one (longer than R regions by the stagger) — must agree with the per-region figure
    span = max(evs[a][0].elapsed_time(evs[b][R]) for a in range(T) for b in range(T))
    SPAN["ms_per_step"] = span / (R * K)
    issue_desc = ("g2048_step_list_timed via StepSchedule from %d host threads on %d streams (thread t issues env sets t, t+%d, "
                  "...; a region is %d launches per stream; one kernel launch per env step); %s"
                  % (T, T, T, Kt, "chained launches (G2048_FLAG_CHAINED | G2048_FLAG_CHAIN_INTERLEAVED)" if chained else "plain launches"))
    return max_over_ranks_vec(ms), issue_desc, t_start, t_end


def rank_cpus(allowed, local, world):
    """The logical CPUs of rank `local`: whole physical cores (hyper-thread siblings stay together, so that no rank's
    issuing thread shares a core with another rank's), dealt out in contiguous runs of cores."""
    groups, seen = [], set()
    for c in allowed:
        if c in seen:
            continue
        sib = {c}
        try:
            with open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c) as f:
                for part in f.read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    sib.update(range(int(lo), int(hi or lo) + 1))
        except (OSError, ValueError):
            pass
        sib = sorted(x for x in sib if x in allowed)
        seen.update(sib)
        groups.append(sib)
    per = len(groups) // world
    if per < 1:
        return []
    return [c for grp in groups[local * per:(local + 1) * per] for c in grp]


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
    if world > 1 and args.pin_cores and hasattr(os, "sched_setaffinity"):
        # N processes x (main + 2 issuing threads + NCCL proxy) on one box: give every rank its own cores, so that no
        # rank's issuing threads wait for a core another rank is spinning on (max over ranks is what is reported)
        try:
            mine = rank_cpus(sorted(os.sched_getaffinity(0)), local, world)
            if len(mine) >= 2:
                os.sched_setaffinity(0, mine)
        except OSError:
            pass
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

    def max_over_ranks_vec(xs):
        t = torch.tensor(xs, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    K, W, R, S = args.steps, args.warmup, args.repeats, args.sets
    lib = g._lib.lib()

    def make_workload(n, total_envs):
        """S env sets of n boards (global ids: set s, rank r -> s*total + r*n ...), boards produced by the env itself
        (reset + 6 real steps), and a pool of uniform-random action rows drawn on the device from the library's
        policy stream — a function of the GLOBAL env id, so the workload does not depend on the sharding."""
        games = [g.BatchedGame2048(n, seed=42, device=dev, env_id_base=s * total_envs + rank * n, outputs=())
                 for s in range(S)]
        pool = torch.empty((args.action_pool, n), dtype=torch.uint8, device=dev)
        for t in range(args.action_pool):
            g._lib.check(lib.g2048_sample_actions(None, pool[t].data_ptr(), n, rank * n, 1234, t,
                                                  torch.cuda.current_stream(dev).cuda_stream))
        for gm in games:
            gm.reset()
            for t in range(6):
                gm.step(pool[t % args.action_pool])
        torch.cuda.synchronize()
        return games, pool

    # ---- headline: BASELINE config 3 at this N (strong: args.envs boards in total; weak: per GPU) ----------
    n = args.envs if args.scaling == "weak" else args.envs // world
    total_envs = n * world
    small = n < args.small_below
    games, pool = make_workload(n, total_envs)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_first_sample()            # BEFORE the spin-up: nothing sleeps between warm-up and the clock
    t_spin = time.perf_counter()
    chained = "interleaved" if args.chain == "on" else False
    T = args.issue_threads
    if T == 0:                       # auto: a second issuing thread once the GPU outruns one (~2.4 us per launch)
        T = 2 if (small and S % 2 == 0 and K % 2 == 0) else 1

    def regions(gm, pl, spin, Kx, Rx, small_x, ch):
        if T > 1 and small_x:
            return time_step_regions_mt(torch, g, gm, pl, spin, W, Kx, Rx, T, barrier, max_over_ranks_vec, ch)
        return time_step_regions(torch, g, gm, pl, spin, W, Kx, Rx, small_x, barrier, max_over_ranks_vec, ch)
    region_ms, issue, t_start, t_end = regions(games, pool, args.spinup, K, R, small, chained)
    headline_span = SPAN["ms_per_step"] if (T > 1 and small) else None
    clocks = sampler.stop(t_spin, t_end) if sampler else None
    if clocks is not None:
        clocks["window"] = "spin-up + warm-up + timed regions (contiguous step launches)"
    ms = statistics.median(region_ms)
    value = total_envs * K / (ms * 1e-3)
    launch_s = ms * 1e-3 / K              # a region holds only step-kernel launches, back to back
    checksum = state_checksum(torch, dist, games, total_envs, rank, n, world, dev)
    # Long regions (`--long-region` launches between two events, median of 5): an event between two launches is a full
    # drain and ramp of the launch pipeline — every region of K launches pays one, ~5 % of a region at the driver's K = 20.
    long_region = None
    if args.long_region > 0:
        Lr = args.long_region - args.long_region % T
        l_ms, _, _, _ = regions(games, pool, max(args.spinup // 2, 256), Lr, 5, small, chained)   # (the clocks ramp again after the checksum's idle gap)
        lm = statistics.median(l_ms)
        long_region = {"launches": Lr, "repeats": len(l_ms), "ms_per_step": lm / Lr, "ms_per_step_max": max(l_ms) / Lr,
                       "value": total_envs * Lr / (lm * 1e-3), "unit": UNIT,
                       "roofline_frac": ALG_BYTES_PER_STEP * n / (lm * 1e-3 / Lr) / 1e9 / hbm_peak()[0]}
    # the same launches without the chain (every launch waits for the whole previous grid): what round 2's first
    # half measured; continues the same env sets, after the checksum
    plain = None
    if chained and not args.no_plain:
        p_ms, p_issue, _, _ = regions(games, pool, max(args.spinup // 2, 256), K, max(R // 2, 5), small, False)
        pm = statistics.median(p_ms)
        plain = {"value": total_envs * K / (pm * 1e-3), "unit": UNIT, "ms_per_step": pm / K, "repeats": len(p_ms),
                 "roofline_frac": ALG_BYTES_PER_STEP * n / (pm * 1e-3 / K) / 1e9 / hbm_peak()[0], "issue": p_issue}

    # ---- weak-scaling extra at N > 1 (args.envs boards PER GPU): what round 1's SCALE file measured ---------
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        del games, pool
        torch.cuda.empty_cache()
        nw = args.envs
        wg, wp = make_workload(nw, nw * world)
        w_ms, w_issue, _, _ = regions(wg, wp, max(args.spinup // 2, 256), K, max(R // 2, 5), nw < args.small_below, chained)
        wm = statistics.median(w_ms)
        weak = {"scaling": "weak", "envs_per_gpu": nw, "global_envs": nw * world, "value": nw * world * K / (wm * 1e-3),
                "unit": UNIT, "ms_per_step": wm / K, "repeats": len(w_ms),
                "roofline_frac": ALG_BYTES_PER_STEP * nw / (wm * 1e-3 / K) / 1e9 / hbm_peak()[0], "issue": w_issue}
        games, pool = wg, wp
        n_f, total_f = nw, nw * world
    else:
        n_f, total_f = n, total_envs

    # ---- several steps per launch (g2048_step_many): the same steps with the boards held in registers in
    #      between — the open-loop form of this workload (the actions are pre-generated).  Reported beside the
    #      headline, never instead of it: a step here moves 6 bytes, not 38.
    fused = None
    if args.fused_steps > 0:
        Kf = args.fused_steps
        gen = torch.Generator(device=dev).manual_seed(99 + rank)
        facts = torch.randint(0, 4, (Kf, n_f), generator=gen, device=dev, dtype=torch.uint8)
        frew = torch.empty((Kf, n_f), dtype=torch.float32, device=dev)
        fdone = torch.empty((Kf, n_f), dtype=torch.uint8, device=dev)
        launches = max(2, min(200, (1 << 31) // (Kf * n_f)))
        for i in range(2):
            games[i % S].step_many(facts, rewards=frew, dones=fdone)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record()
        for i in range(launches):
            games[i % S].step_many(facts, rewards=frew, dones=fdone)
        f1.record()
        torch.cuda.synchronize()
        f_ms = max_over_ranks(f0.elapsed_time(f1))
        fused = {"api": "g2048_step_many", "steps_per_launch": Kf, "launches": launches, "envs_per_gpu": n_f,
                 "value": total_f * Kf * launches / (f_ms * 1e-3), "unit": UNIT,
                 "us_per_step": f_ms * 1e3 / (Kf * launches),
                 "algorithmic_bytes_per_step": 6 + 32.0 / Kf,
                 "note": "bit-identical to steps_per_launch calls of step(); boards stay in registers between steps"}
        del facts, frew, fdone
    del games, pool
    torch.cuda.empty_cache()

    # ---- BASELINE config 4 (N = 1 only): 262,144 envs returning the legal-action mask, auto-reset, actions uniform
    #      among the legal moves — mid-game boards.  One launch per step: the step kernel draws the action from the
    #      mask the previous step wrote, plays it and writes the new mask (step(policy="legal")).
    config4 = None
    if world == 1 and not args.no_config4:
        n4, S4 = 262144, 16
        g4 = [g.BatchedGame2048(n4, seed=42, device=dev, env_id_base=s * n4, outputs=("legal_mask",)) for s in range(S4)]
        for gm in g4:
            gm.reset()
            gm.step_many(policy="legal", n_steps=200)               # play into the mid-game (mean ~5 empty cells)
        # (the kernel WRITES the actions it draws: one row per env set — set s always gets row s, S4 rows in the pool)
        acts4 = torch.empty((S4, n4), dtype=torch.uint8, device=dev)
        K4, R4, spin4 = 50, 25, 4096
        T4 = T if (S4 % T == 0 and K4 % T == 0) else 1
        if T4 > 1:
            r4, issue4, _, _ = time_step_regions_mt(torch, g, g4, acts4, spin4, 0, K4, R4, T4, barrier, max_over_ranks_vec,
                                                    chained, policy="legal")
        else:
            r4, issue4, _, _ = time_step_regions(torch, g, g4, acts4, spin4, 0, K4, R4, True, barrier, max_over_ranks_vec,
                                                 chained, policy="legal")
        m4 = statistics.median(r4) / K4                                                      # ms per step
        empties = float((g4[0].boards == 0).sum(dim=1).float().mean())
        bytes4 = 40                                   # board in 16 + mask in 1 + board out 16 + mask out 1 + action 1 + reward 4 + done 1
        peak4 = hbm_peak()[0]
        config4 = {"workload": "BASELINE config 4: %d envs, legal-action mask + auto-reset, actions uniform among the "
                               "legal moves (drawn in the step kernel), %d env sets round-robin" % (n4, S4),
                   "value": n4 / (m4 * 1e-3), "unit": UNIT, "us_per_step": m4 * 1e3, "launches_per_step": 1,
                   "kernel": "g2048_step_kernel<O_MASK, false, POLICY_LEGAL>", "algorithmic_bytes_per_step": bytes4,
                   "roofline_frac": bytes4 * n4 / (m4 * 1e-3) / 1e9 / peak4, "mean_empty_cells": empties,
                   "issue": issue4, "steps": K4, "repeats": R4}
        del g4, acts4
        torch.cuda.empty_cache()

    # ---- e2e: host buffers through the C-ABI handle, copies inside the timed region ----------
    # The caller hands over pinned HOST actions and gets boards [n,16] / rewards / dones in HOST arrays, every step.
    #   e2e            the handle's default: the boards cross PCIe 4 bits per cell and the library's host threads expand
    #                  them into the caller's [n,16] array inside the call (G2048_BOARDS_BYTES_PACKED_WIRE)
    #   e2e_plain_wire the same call moving the 16 bytes per board as they are (round 1 / first half of round 2)
    #   e2e_compact    the caller ASKS for the packed boards ([n,8]; nothing is expanded)
    Ke = max(3, min(K * R, args.e2e_steps))
    host_pool = torch.randint(0, 4, (8, n), dtype=torch.uint8).pin_memory()

    def time_host_env(**kw):
        henv = g.HostSteppedEnv(n, seed=42, device=local, env_id_base=rank * n, **kw)
        henv.reset()
        for i in range(5):
            henv.step_pinned(host_pool[i % 8])
        barrier()
        t0 = time.perf_counter()
        overflow = 0
        for i in range(Ke):
            overflow += henv.step_pinned(host_pool[i % 8]).nibble_overflow.value
        secs = max_over_ranks(time.perf_counter() - t0)
        chk = float(henv.buffers.rewards.sum())    # the result is read on the host
        bsum = int(henv.buffers.boards.to(torch.int64).sum()) if kw.get("board_format") != "nibble" else None
        launches = henv.n_chunks_effective * Ke
        henv.close()
        return secs, chk, bsum, overflow, launches
    e2e_s, chk, bsum, e2e_overflow, e2e_launches = time_host_env(n_chunks=args.e2e_chunks_packed, wire="packed",
                                                                 unpack_threads=args.unpack_threads)
    e2e_value = total_envs * Ke / e2e_s
    ep_s, chk_p, bsum_p, _, _ = time_host_env(n_chunks=args.e2e_chunks, wire="plain")
    e2e_plain = {"value": total_envs * Ke / ep_s, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": n * 21,
                 "steps": Ke, "ms_per_step": 1e3 * ep_s / Ke, "chunks": args.e2e_chunks, "checksum": chk_p,
                 "boards_checksum": bsum_p,
                 "api": "g2048_env_step_host, G2048_BOARDS_BYTES (16 bytes per board over PCIe, pinned host buffers)"}
    e2c_s, chk_c, _, overflow, _ = time_host_env(n_chunks=args.e2e_chunks, board_format="nibble")
    e2e_compact = {"value": total_envs * Ke / e2c_s, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": n * 13,
                   "steps": Ke, "ms_per_step": 1e3 * e2c_s / Ke, "boards_not_fitting": overflow, "checksum": chk_c,
                   "api": "g2048_env_step_host, G2048_BOARDS_NIBBLE (boards 4 bits per cell, pinned host buffers)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = hbm_peak()
    achieved = ALG_BYTES_PER_STEP * n / launch_s / 1e9
    traffic, traffic_src, traffic_kind = committed_ncu_traffic(n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "repeats": R,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
        "timing": {"issue": issue, "statistic": "median over `repeats` back-to-back regions of `steps` launches (CUDA events on the "
                                "launching stream, max over ranks per region)",
                   "issue_threads": T, "ms_per_step_all_regions_span": headline_span,
                   "spinup_launches": args.spinup, "ms_per_step_min": min(region_ms) / K,
                   "ms_per_step_max": max(region_ms) / K, "ms_per_step_mean": statistics.fmean(region_ms) / K},
        "state_checksum": checksum,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_definition": "algorithmic bytes (38 B per board-step, SURVEY 8d) / launch time / peak",
                     "traffic": traffic, "traffic_source": traffic_src, "traffic_kind": traffic_kind,
                     "peak_source": peak_src, "kernel": "g2048_step_kernel<0, false>",
                     "algorithmic_bytes_per_launch": ALG_BYTES_PER_STEP * n,
                     "launch_us": launch_s * 1e6, "frac_of_nominal_8TBs": achieved / 8000.0,
                     "launch_time_definition": "region time / launches in the region (a region holds nothing but step launches). "
                                               "Chained launches of different env sets overlap on the GPU, so this is the rate at "
                                               "which launches complete; one chained launch running ALONE (as under ncu, which "
                                               "serialises launches) takes longer — profiles/r02_chain_launches_summary.md"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": n * 13,
                "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke, "chunks": args.e2e_chunks_packed or 4,
                "api": "g2048_env_step_host, G2048_BOARDS_BYTES_PACKED_WIRE: pinned host buffers, boards [n,16] exponent bytes "
                       "in the caller's array; on the wire 4 bits per cell, expanded inside the call by the library's host "
                       "threads", "unpack_threads": args.unpack_threads or "library default (half the CPUs of the process, at most 8)",
                "boards_not_fitting": e2e_overflow, "checksum": chk, "boards_checksum": bsum},
        "e2e_plain_wire": e2e_plain,
        "e2e_compact": e2e_compact,
        "weak": weak,
        "plain_launches": plain,
        "long_region": long_region,
        "fused": fused,
        "config4": config4,
        "gpu_launches": K * R, "e2e_gpu_launches": e2e_launches,
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
    ap.add_argument("--steps", type=int, default=None, help="K: step launches per timed region")
    ap.add_argument("--warmup", type=int, default=None, help="W: untimed launches right before the first region")
    ap.add_argument("--repeats", type=int, default=25, help="R: back-to-back timed regions; the median is reported")
    ap.add_argument("--spinup", type=int, default=16384,
                    help="untimed step launches before the warm-up (clock/power ramp; fixed, so the state "
                         "checksum does not depend on N)")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=1 << 20, help="boards in total (strong, BASELINE config 3) or per GPU (weak)")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling extra block at N > 1")
    ap.add_argument("--sets", type=int, default=32)
    ap.add_argument("--small-below", type=int, default=1 << 21,
                    help="shards smaller than this are issued through g2048_step_list (K launches per C call): a "
                         "Python loop issues a launch every ~6 us, a B200 steps 524,288 boards in ~4.7 us and "
                         "131,072 in ~1.3 us (profiles/r02_launch_rate.log); 0 forces the Python loop")
    ap.add_argument("--action-pool", type=int, default=16)
    ap.add_argument("--fused-steps", type=int, default=32, help="steps per g2048_step_many launch (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--e2e-chunks", type=int, default=2)
    ap.add_argument("--e2e-chunks-packed", type=int, default=0, help="slices of the default (packed-wire) host call; 0 = library default")
    ap.add_argument("--unpack-threads", type=int, default=0, help="host threads of the packed-wire host call; 0 = library default")
    ap.add_argument("--ref-steps-per-proc", type=int, default=5000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the BASELINE config 4 extra block (N = 1)")
    ap.add_argument("--chain", choices=["on", "off"], default="on",
                    help="on: the step launches are chained launches (include/g2048.h, G2048_FLAG_CHAINED): each depends "
                         "on the previous step of its own env set warp by warp; off: plain launches")
    ap.add_argument("--issue-threads", type=int, default=0,
                    help="host threads (and streams) that issue the step launches through g2048_step_list, each its own "
                         "env sets; 0 = auto (2 when --sets and --steps are even: one thread issues a launch every ~2.4 us, "
                         "two ~1.4 us between them, a B200 steps 131,072 boards in ~1.3 us; profiles/r02_issue_threads.log)")
    ap.add_argument("--pin-cores", type=int, default=1,
                    help="N > 1: restrict every rank to its own 1/N of the host cores (1) or leave placement to the OS (0)")
    ap.add_argument("--long-region", type=int, default=1000, help="launches of a long timed region (extra block, median of 5; 0 = skip)")
    ap.add_argument("--no-plain", action="store_true", help="skip the plain-launch extra block (the headline's launches unchained)")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        args.steps = 20 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 200 if args.steps is None else args.steps
    args.warmup = 50 if args.warmup is None else max(args.warmup, 3)
    args.repeats = max(args.repeats, 1)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
