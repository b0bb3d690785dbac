"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous sharding by global env id, no
collective on the step path, and the optional all-reduce of episode statistics.  Each rank steps
its slice with the C oracle (the CUDA path needs a GPU; its sharding invariance is covered by
tests/test_gpu_parity.py) and the reduced statistics must equal the unsharded run's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gym_2048_b200 as g
from oracle import oracle

N, STEPS, SEED = 3001, 40, 17


def _run(n, base, seed, steps, acts):
    env = oracle.OracleBatch(n, seed=seed, env_id_base=base)
    env.reset()
    st = g.EpisodeStats("cpu")
    boards = []
    for t in range(steps):
        o = env.step(acts[t, base:base + n])
        st.update(torch.from_numpy(o["dones"]), torch.from_numpy(o["final_score"].astype(np.int64)),
                  torch.from_numpy(o["final_len"].astype(np.int64)), torch.from_numpy(o["highest_exp"]),
                  torch.from_numpy(o["illegal"]))
        boards.append(env.boards.copy())
    return st, np.stack(boards)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acts = np.random.default_rng(5).integers(0, 4, (STEPS, N)).astype(np.uint8)
    base, count = g.shard_range(N, rank, world)
    st, boards = _run(count, base, SEED, STEPS, acts)
    st.all_reduce()
    np.save(os.path.join(out_dir, "boards_%d.npy" % rank), boards)
    if rank == 0:
        torch.save(st.vec, os.path.join(out_dir, "stats.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_rollout_equals_unsharded(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    acts = np.random.default_rng(5).integers(0, 4, (STEPS, N)).astype(np.uint8)
    full_stats, full_boards = _run(N, 0, SEED, STEPS, acts)
    sharded = np.concatenate([np.load(tmp_path / ("boards_%d.npy" % r)) for r in range(2)], axis=1)
    assert np.array_equal(sharded, full_boards)                       # no data-path exchange needed
    reduced = torch.load(tmp_path / "stats.pt")
    assert torch.equal(reduced, full_stats.vec)
    s = full_stats.summary()
    assert s["episodes"] > 100 and s["mean_length"] > 5 and sum(s["highest_tile_hist"].values()) == s["episodes"]


# ---- bench.py's state checksum (the proof of sharding invariance the SCALE run prints) on two gloo ranks -----------
class _Set:
    """What bench.state_checksum reads of a BatchedGame2048: `.boards`, uint8 [n,16]."""

    def __init__(self, boards):
        self.boards = torch.from_numpy(np.ascontiguousarray(boards))


def _checksum_worker(rank, world, port, out_dir):
    import importlib.util
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    total, sets = 4096, 3
    n = total // world
    games = []
    for s in range(sets):                                   # bench.py's id layout: set s, rank r -> s*total + r*n
        env = oracle.OracleBatch(n, seed=42, env_id_base=s * total + rank * n)
        env.reset()
        acts = np.random.default_rng(100 + s).integers(0, 4, (10, total)).astype(np.uint8)
        for t in range(10):
            env.step(acts[t, rank * n:(rank + 1) * n])
        games.append(_Set(env.boards))
    c = bench.state_checksum(torch, dist, games, total, rank, n, world, torch.device("cpu"))
    if rank == 0:
        open(os.path.join(out_dir, "checksum_%d.txt" % world), "w").write(c)
    dist.barrier()
    dist.destroy_process_group()


def test_bench_state_checksum_is_the_same_for_one_and_two_ranks(tmp_path):
    """The 64-bit state checksum bench.py prints (all boards of all env sets, keyed by the global env id, summed over
    the ranks) does not depend on the number of ranks the batch is sharded over — and it does depend on the boards."""
    sums = {}
    for world in (1, 2):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        mp.spawn(_checksum_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
        sums[world] = open(tmp_path / ("checksum_%d.txt" % world)).read()
    assert sums[1] == sums[2] and len(sums[1]) == 16
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = np.zeros((64, 16), np.uint8)
    b = a.copy()
    b[5, 3] = 1
    sw = a.copy()
    sw[[0, 1]] = [[1] + [0] * 15, [2] + [0] * 15]
    sw2 = a.copy()
    sw2[[0, 1]] = [[2] + [0] * 15, [1] + [0] * 15]           # the same boards on swapped env ids: a different state
    cs = [bench.state_checksum(torch, dist, [_Set(x)], 64, 0, 64, 1, torch.device("cpu")) for x in (a, b, sw, sw2)]
    assert len(set(cs)) == 4
