"""GPU parity tests proper: the product (libg2048.so through its C ABI, driven by
gym_2048_b200.BatchedGame2048) against the golden vectors produced by the unmodified
reference and against the pinned C oracle, bit for bit."""
import numpy as np
import pytest

import parity_checks as pc
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from backends import GpuOps
    return GpuOps


def test_philox_on_device(ops):
    pc.check_philox(ops)


def test_draw_words_on_device(ops):
    pc.check_draw_words(ops)


def test_shift_table_all_directions(ops):
    pc.check_shift_table_all_directions(ops)


def test_csv_transitions(ops):
    pc.check_csv_transitions(ops)


def test_special_boards(ops):
    pc.check_special_boards(ops)


def test_golden_rollouts(ops, golden_rollouts):
    pc.check_rollouts(ops, golden_rollouts)


def test_status_and_move_random_boards(ops):
    pc.check_status_random_boards(ops)


def test_add_tile(ops):
    pc.check_add_tile(ops)


def test_config2_65536_envs_fixture_seeds(ops):
    """BASELINE config 2: 65,536 envs, random actions, bit-exact vs CPU on the fixture seeds."""
    for seed in (0, 1, 42, 456):
        pc.check_against_oracle(ops, n=65536, steps=24, seed=seed, policy="random", threads=8)


def test_config2_long_random_rollout(ops):
    pc.check_against_oracle(ops, n=65536, steps=256, seed=0, policy="random", threads=8)


def test_config4_legal_mask_autoreset_midgame(ops):
    """BASELINE config 4 flavour: legal-mask output + auto-reset, random-legal play (mid-game boards)."""
    pc.check_against_oracle(ops, n=16384, steps=600, seed=42, policy="legal", illegal_move_reward=-1.0,
                            env_id_base=(1 << 40) + 3, threads=8)


def test_max_tile_and_no_autoreset(ops):
    pc.check_against_oracle(ops, n=8192, steps=300, seed=456, policy="legal", max_tile_exp=6, auto_reset=False,
                            threads=8)
    pc.check_against_oracle(ops, n=8192, steps=300, seed=1, policy="legal", max_tile_exp=7, auto_reset=True,
                            threads=8)


def test_ragged_and_tiny_batches(ops):
    for n in (1, 2, 31, 33, 255, 257, 1000):
        pc.check_against_oracle(ops, n=n, steps=40, seed=n, policy="random", threads=1)


def test_sharding_invariance_on_device(ops):
    """Slices with env_id_base offsets reproduce the unsharded batch (no collective on the path)."""
    import torch
    import gym_2048_b200 as g
    n, steps = 8192, 50
    rng = np.random.default_rng(9)
    acts = rng.integers(0, 4, (steps, n)).astype(np.uint8)
    full = g.BatchedGame2048(n, seed=11, env_id_base=100, outputs=())
    parts = [g.BatchedGame2048(c, seed=11, env_id_base=100 + b, outputs=())
             for b, c in (g.shard_range(n, r, 3) for r in range(3))]
    full.reset()
    for p in parts:
        p.reset()
    for t in range(steps):
        a = torch.from_numpy(acts[t]).cuda()
        rf = full.step(a)
        lo = 0
        for p in parts:
            rp = p.step(a[lo:lo + p.num_envs].contiguous())
            assert torch.equal(rp.boards, rf.boards[lo:lo + p.num_envs])
            assert torch.equal(rp.rewards, rf.rewards[lo:lo + p.num_envs])
            assert torch.equal(rp.dones, rf.dones[lo:lo + p.num_envs])
            lo += p.num_envs


def test_lean_kernel_equals_full_kernel(ops):
    """outputs=() runs the lean kernel variant; state, rewards and dones must not differ."""
    import torch
    import gym_2048_b200 as g
    n = 50000
    a = g.BatchedGame2048(n, seed=3, outputs=())
    b = g.BatchedGame2048(n, seed=3)
    a.reset(), b.reset()
    gen = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(100):
        act = torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.uint8)
        ra, rb = a.step(act), b.step(act)
        assert torch.equal(ra.boards, rb.boards) and torch.equal(ra.rewards, rb.rewards)
        assert torch.equal(ra.dones, rb.dones)


def test_full_size_properties_1m():
    """BASELINE config 3 size (1,048,576 envs): size-independent properties + a sampled oracle check."""
    import torch
    import gym_2048_b200 as g
    n = 1 << 20
    game = g.BatchedGame2048(n, seed=42)
    game.reset()
    b0 = game.boards.clone()
    assert int((b0 != 0).sum()) == 2 * n                      # two tiles per fresh board (:108-109)
    assert int(b0.max()) <= 2
    gen = torch.Generator(device="cuda").manual_seed(5)
    sample = torch.randint(0, n, (4096,), generator=gen, device="cuda")
    for t in range(20):
        before = game.boards.clone()
        tiles_before = game.board_values().sum(dim=(1, 2))
        act = torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.uint8)
        r = game.step(act)
        legal = ~r.illegal
        # an illegal move leaves the terminal board untouched and pays the illegal reward
        assert torch.equal(r.terminal_boards[r.illegal], before[r.illegal])
        assert bool((r.rewards[r.illegal] == 0).all())
        # tile-value conservation: a legal move adds exactly the spawned 2 or 4
        post = torch.where(r.dones[:, None], r.terminal_boards, r.boards)
        grow = game.board_values(post.contiguous()).sum(dim=(1, 2)) - tiles_before
        assert bool(((grow[legal] == 2) | (grow[legal] == 4)).all()) and bool((grow[r.illegal] == 0).all())
        # every reset board has exactly two tiles
        assert bool(((r.boards[r.dones] != 0).sum(dim=1) == 2).all())
        # sampled envs against the oracle on the same boards/actions/draws
        idx = sample.cpu().numpy()
        ob = oracle.OracleBatch(1, seed=42)
        for i in idx[:256]:
            ob.boards[0] = before[i].cpu().numpy()
            ob.env_id_base, ob.step_index = int(i), t
            o = ob.step(np.array([int(act[i])], dtype=np.uint8))
            assert np.array_equal(ob.boards[0], r.boards[i].cpu().numpy())
            assert o["rewards"][0] == float(r.rewards[i]) and o["dones"][0] == int(r.dones[i])
    # mean P(4) over spawns ~ 0.1
    fours = (b0 == 2).sum().item() / (2 * n)
    assert abs(fours - 0.1) < 0.002


def test_batch_crossing_a_2_32_env_id_boundary(ops):
    """The kernel folds the high half of the env id into a launch-uniform Philox head; a batch
    whose ids cross a multiple of 2^32 is split into two launches by the library."""
    pc.check_against_oracle(ops, n=5000, steps=40, seed=3, policy="random", env_id_base=(1 << 32) - 1234)
    pc.check_against_oracle(ops, n=3000, steps=40, seed=4, policy="legal", env_id_base=(5 << 32) - 1)


def test_lean_kernel_crossing_boundary_and_device_step_counter():
    """Lean outputs (boards, rewards, dones only), ids crossing 2^32, step index kept on the device
    and the loop replayed as a CUDA graph: same boards as host-indexed stepping and as the oracle."""
    import torch
    import gym_2048_b200 as g
    n, base, T = 4096, (1 << 32) - 777, 24
    dev = torch.device("cuda", 0)
    acts = torch.randint(0, 4, (T, n), device=dev, dtype=torch.uint8,
                         generator=torch.Generator(device=dev).manual_seed(8))
    ref = oracle.OracleBatch(n, seed=9, env_id_base=base)
    ref.reset()
    host = g.BatchedGame2048(n, seed=9, env_id_base=base, outputs=())
    host.reset()
    for t in range(T):
        ref.step(acts[t].cpu().numpy())
        host.step(acts[t])
    assert np.array_equal(host.boards.cpu().numpy(), ref.boards)
    # device-side counter, first 8 steps eagerly, then 2 replays of an 8-step graph
    gm = g.BatchedGame2048(n, seed=9, env_id_base=base, outputs=())
    gm.reset()
    gm.use_device_step_counter(True)
    for t in range(8):
        gm.step(acts[t])
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    staged = acts[8:16].clone()
    with torch.cuda.stream(stream):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for t in range(8):
                gm.step(staged[t])
    # capture does not execute: the counter still says 8
    graph.replay()
    torch.cuda.synchronize()
    staged.copy_(acts[16:24])
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(gm.boards.cpu().numpy(), ref.boards)
    gm.use_device_step_counter(False)
    assert gm.step_index == T


@pytest.mark.parametrize("n,K,base,step0,kw", [
    (1, 5, 0, 0, {}),
    (33, 1, 7, 3, {}),
    (1000, 40, 123, 0, dict(illegal_move_reward=-1.0)),
    (5000, 24, 0, 0, dict(auto_reset=False)),
    (4096, 64, 0, 11, dict(max_tile_exp=5)),
    (1000, 6, 2**32 - 500, 0, {}),                    # env ids cross a multiple of 2^32 inside the call
    (700, 7, 5, 2**32 - 3, {}),                       # step indices cross a multiple of 2^32 inside the call
    (70001, 16, 2**33 + 1, 2**40, {}),                # several boards per thread, ragged tail
])
def test_step_many_is_k_steps(n, K, base, step0, kw):
    """g2048_step_many == K calls of step(): against the oracle stepped K times (rewards, dones, illegal, the
    board after every step) and against the product's own single-step kernel."""
    import torch
    import gym_2048_b200 as g
    from oracle import oracle
    seed = 1234567
    rng = np.random.default_rng(n + K)
    actions = rng.integers(0, 4, (K, n)).astype(np.uint8)
    max_tile = None if not kw.get("max_tile_exp") else 2 ** kw["max_tile_exp"]
    mk = lambda outputs: g.BatchedGame2048(n, seed=seed, env_id_base=base, outputs=outputs, max_tile=max_tile,
                                           illegal_move_reward=kw.get("illegal_move_reward", 0.0),
                                           auto_reset=kw.get("auto_reset", True))
    ref = oracle.OracleBatch(n, seed=seed, env_id_base=base, **kw)
    ref.reset()
    ref.step_index = step0
    many, single, lean = mk(("illegal",)), mk(("illegal",)), mk(())
    for game in (many, single, lean):
        game.reset()
        game.step_index = step0
        assert np.array_equal(game.boards.cpu().numpy(), ref.boards)
    d_act = torch.from_numpy(actions).cuda()
    illegal = torch.empty((K, n), dtype=torch.uint8, device="cuda")
    traj = torch.empty((K, n, 16), dtype=torch.uint8, device="cuda")
    rewards, dones = many.step_many(d_act, illegal=illegal, boards_traj=traj)
    rewards_lean, dones_lean = lean.step_many(d_act)
    assert many.step_index == step0 + K
    for k in range(K):
        o = ref.step(actions[k])
        r = single.step(d_act[k])
        assert np.array_equal(rewards[k].cpu().numpy(), o["rewards"]), k
        assert np.array_equal(dones[k].cpu().numpy().astype(np.uint8), o["dones"]), k
        assert np.array_equal(illegal[k].cpu().numpy(), o["illegal"]), k
        assert np.array_equal(traj[k].cpu().numpy(), ref.boards), k
        assert torch.equal(r.rewards, rewards[k]) and torch.equal(r.dones, dones[k]) and torch.equal(r.boards, traj[k])
    assert torch.equal(many.boards, traj[K - 1]) and torch.equal(lean.boards, many.boards)
    assert torch.equal(rewards_lean, rewards) and torch.equal(dones_lean, dones)


@pytest.mark.parametrize("policy", ["uniform", "legal"])
@pytest.mark.parametrize("n,K,base,step0,kw", [
    (1000, 48, 77, 5, dict(illegal_move_reward=-1.0)),
    (4096, 200, 0, 0, {}),                            # random-legal play reaches mid-game boards and natural deaths
    (3000, 12, 2**32 - 1500, 2**32 - 4, dict(auto_reset=False)),
])
def test_step_many_device_policy_is_sample_actions_plus_step(policy, n, K, base, step0, kw):
    """g2048_step_many with a policy flag == sample_actions() followed by step(), K times: against the oracle's
    sampler and step (actions, rewards, dones, boards, legal masks) and against the product's two-kernel loop."""
    import torch
    import gym_2048_b200 as g
    from oracle import oracle
    seed = 99
    legal = policy == "legal"
    mk = lambda: g.BatchedGame2048(n, seed=seed, env_id_base=base, outputs=("legal_mask",),
                                   illegal_move_reward=kw.get("illegal_move_reward", 0.0),
                                   auto_reset=kw.get("auto_reset", True))
    ref = oracle.OracleBatch(n, seed=seed, env_id_base=base, **kw)
    ref.reset()
    ref.step_index = step0
    fused, loop = mk(), mk()
    for game in (fused, loop):
        game.reset()
        game.step_index = step0
    acts = torch.empty((K, n), dtype=torch.uint8, device="cuda")
    masks = torch.empty((K, n), dtype=torch.uint8, device="cuda")
    traj = torch.empty((K, n, 16), dtype=torch.uint8, device="cuda")
    rewards, dones = fused.step_many(policy=policy, n_steps=K, actions_out=acts, legal_mask_out=masks, boards_traj=traj)
    mask = oracle.status(ref.boards)["legal_mask"]
    for k in range(K):
        want = oracle.sample_actions(mask if legal else None, n, base, seed, step0 + k)
        assert np.array_equal(acts[k].cpu().numpy(), want), k
        o = ref.step(want)
        mask = o["legal_mask"]
        assert np.array_equal(rewards[k].cpu().numpy(), o["rewards"]), k
        assert np.array_equal(dones[k].cpu().numpy().astype(np.uint8), o["dones"]), k
        assert np.array_equal(traj[k].cpu().numpy(), ref.boards), k
        assert np.array_equal(masks[k].cpu().numpy(), mask), k
        a = loop.sample_actions(legal=legal)
        assert torch.equal(a, acts[k])
        r = loop.step(a)
        assert torch.equal(r.boards, traj[k]) and torch.equal(r.rewards, rewards[k])
    assert torch.equal(fused.boards, loop.boards) and torch.equal(fused.legal_mask, loop.legal_mask)


# ---------------------------------------------------------------------------------------------------------
# round 2: specialised output sets, g2048_step_n, graph capture, real multi-GPU sharding
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("outputs", [
    (),                                                   # lean kernel
    ("legal_mask",),                                      # O_MASK kernel (BASELINE config 4)
    ("illegal", "highest", "legal_mask"),                 # the evaluator's kernel
    ("illegal", "highest", "legal_mask", "episode", "terminal"),   # Game2048VecEnv's kernel
    ("illegal",), ("episode",), ("highest", "terminal"),  # odd combinations: the generic (nullable-pointer) kernel
])
def test_every_output_set_matches_the_oracle(outputs):
    """Each compile-time output set the library instantiates (and the generic kernel behind any other
    combination) against the oracle: boards, rewards, dones always; the optional outputs that were asked for."""
    import torch
    import gym_2048_b200 as g
    n, T = 70001, 80                                       # ragged: partial last warp, several boards per thread at 1024x148? no: small shape
    gm = g.BatchedGame2048(n, seed=21, env_id_base=5, outputs=outputs, illegal_move_reward=-1.0, max_tile=64)
    ref = oracle.OracleBatch(n, seed=21, env_id_base=5, illegal_move_reward=-1.0, max_tile_exp=6, threads=4)
    assert np.array_equal(gm.reset().cpu().numpy(), ref.reset())
    rng = np.random.default_rng(2)
    for t in range(T):
        act = rng.integers(0, 4, n).astype(np.uint8)
        r = gm.step(torch.from_numpy(act).cuda())
        o = ref.step(act)
        d = o["dones"] != 0
        assert np.array_equal(r.boards.cpu().numpy(), ref.boards), t
        assert np.array_equal(r.rewards.cpu().numpy(), o["rewards"]) and np.array_equal(r.dones.cpu().numpy(), d)
        if "illegal" in outputs:
            assert np.array_equal(r.illegal.cpu().numpy(), o["illegal"] != 0)
        if "highest" in outputs:
            assert np.array_equal(r.highest_exp.cpu().numpy(), o["highest_exp"])
        if "legal_mask" in outputs:
            assert np.array_equal(r.legal_mask.cpu().numpy(), o["legal_mask"])
        if "terminal" in outputs:
            assert np.array_equal(r.terminal_boards.cpu().numpy()[d], o["terminal_boards"][d])
        if "episode" in outputs:
            assert np.array_equal(gm.ep_score.cpu().numpy().astype(np.uint32), ref.ep_score)
            assert np.array_equal(gm.ep_len.cpu().numpy().astype(np.uint32), ref.ep_len)
            assert np.array_equal(gm.ep_return.cpu().numpy(), ref.ep_return)
            assert np.array_equal(r.final_score.cpu().numpy().astype(np.uint32)[d], o["final_score"][d])
            assert np.array_equal(r.final_len.cpu().numpy().astype(np.uint32)[d], o["final_len"][d])
            assert np.array_equal(r.final_return.cpu().numpy()[d], o["final_return"][d])


@pytest.mark.parametrize("n,K,outputs", [
    (1, 3, ()), (1000, 17, ("legal_mask",)), (131072, 24, ()), (70001, 9, ("illegal", "highest", "legal_mask", "episode")),
])
def test_step_n_is_k_calls_of_step(n, K, outputs):
    """g2048_step_n (K launches issued from C) == K calls of step(): per-step rewards/dones/optional rows, final
    boards, running episode statistics — against the product's own step() and, through it, the oracle."""
    import torch
    import gym_2048_b200 as g
    mk = lambda: g.BatchedGame2048(n, seed=4, env_id_base=2**32 - 500, outputs=outputs, illegal_move_reward=-1.0)   # noqa: E731
    a, b = mk(), mk()
    a.reset(), b.reset()
    acts = torch.randint(0, 4, (K, n), device="cuda", dtype=torch.uint8, generator=torch.Generator(device="cuda").manual_seed(1))
    ill = torch.zeros((K, n), dtype=torch.uint8, device="cuda") if "illegal" in outputs else None
    hi = torch.zeros((K, n), dtype=torch.uint8, device="cuda") if "highest" in outputs else None
    lm = torch.zeros((K, n), dtype=torch.uint8, device="cuda") if "legal_mask" in outputs else None
    rew, done = a.step_n(acts, illegal=ill, highest_exp=hi, legal_mask_out=lm)
    for k in range(K):
        r = b.step(acts[k])
        assert torch.equal(rew[k], r.rewards) and torch.equal(done[k], r.dones), k
        if ill is not None:
            assert torch.equal(ill[k].bool(), r.illegal)
        if hi is not None:
            assert torch.equal(hi[k], r.highest_exp)
        if lm is not None:
            assert torch.equal(lm[k], r.legal_mask)
    assert torch.equal(a.boards, b.boards) and a.step_index == b.step_index == K
    if "legal_mask" in outputs:
        assert torch.equal(a.legal_mask, b.legal_mask)
    if "episode" in outputs:
        assert torch.equal(a.ep_score, b.ep_score) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_return, b.ep_return)
    ref = oracle.OracleBatch(n, seed=4, env_id_base=2**32 - 500, illegal_move_reward=-1.0, threads=4)
    ref.reset()
    for k in range(K):
        ref.step(acts[k].cpu().numpy())
    assert np.array_equal(a.boards.cpu().numpy(), ref.boards)


def test_capture_replays_a_closed_loop_of_policy_and_step():
    """BatchedGame2048.capture(): T x (device policy -> step) captured once, replayed R times == the same loop run
    eagerly on a twin env (and the oracle): the device-side step index makes every replay draw fresh tiles."""
    import torch
    import gym_2048_b200 as g
    n, T, R = 20000, 6, 5
    mk = lambda: g.BatchedGame2048(n, seed=77, outputs=("legal_mask",))     # noqa: E731
    a, b = mk(), mk()
    a.reset(), b.reset()
    acts = torch.zeros(n, dtype=torch.uint8, device="cuda")
    shift = torch.zeros(1, dtype=torch.uint8, device="cuda")

    def policy(game, out):
        # a deterministic closed-loop policy on the device: lowest legal move, rotated by a replay counter
        m = game.legal_mask
        first = torch.where(m & 1 != 0, 0, torch.where(m & 2 != 0, 1, torch.where(m & 4 != 0, 2, 3)))
        out.copy_(((first + shift) & 3).to(torch.uint8))

    def body():
        for _ in range(T):
            policy(a, acts)
            a.step(acts)
    replay = a.capture(body, warmup=1)
    assert replay.steps == T
    for r in range(R):
        shift.fill_(r & 3)
        replay()
    torch.cuda.synchronize()
    a.use_device_step_counter(False)
    assert a.step_index == T * (R + 1)
    ref = oracle.OracleBatch(n, seed=77)
    ref.reset()
    acts_b = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for r in range(-1, R):
        shift.fill_(max(r, 0) & 3 if r >= 0 else 0)
        for _ in range(T):
            policy(b, acts_b)
            b.step(acts_b)
            ref.step(acts_b.cpu().numpy())
    assert torch.equal(a.boards, b.boards)
    assert np.array_equal(a.boards.cpu().numpy(), ref.boards)


def test_sharding_invariance_on_two_real_gpus():
    """SURVEY §8e on hardware: the same 1 Mi-id batch stepped as one batch on cuda:0 and as two halves on cuda:0 /
    cuda:1 (env_id_base = rank * n/2) gives bit-identical boards, rewards and dones."""
    import torch
    import gym_2048_b200 as g
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, T = 1 << 18, 40
    full = g.BatchedGame2048(n, seed=13, device="cuda:0", env_id_base=1000, outputs=("legal_mask",))
    halves = [g.BatchedGame2048(n // 2, seed=13, device="cuda:%d" % r, env_id_base=1000 + r * (n // 2), outputs=("legal_mask",))
              for r in range(2)]
    full.reset()
    for h in halves:
        h.reset()
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    for t in range(T):
        act = torch.randint(0, 4, (n,), device="cuda:0", dtype=torch.uint8, generator=gen)
        rf = full.step(act)
        rs = [h.step(act[r * (n // 2):(r + 1) * (n // 2)].to(h.device)) for r, h in enumerate(halves)]
        for key in ("boards", "rewards", "dones", "legal_mask"):
            got = torch.cat([getattr(x, key).to("cuda:0") for x in rs])
            assert torch.equal(getattr(rf, key), got), (t, key)


def test_step_schedule_equals_eager_round_robin():
    """StepSchedule / g2048_step_list: K pre-built step calls over three env sets, issued by one C call in two
    slices, == the same calls made one by one from Python."""
    import torch
    import gym_2048_b200 as g
    n, K = 33000, 30
    mk = lambda s: g.BatchedGame2048(n, seed=3, env_id_base=s * n, outputs=("legal_mask",) if s == 1 else ())   # noqa: E731
    A, B = [mk(s) for s in range(3)], [mk(s) for s in range(3)]
    for x in A + B:
        x.reset()
    acts = torch.randint(0, 4, (K, n), device="cuda", dtype=torch.uint8, generator=torch.Generator(device="cuda").manual_seed(2))
    sched = g.StepSchedule()
    for j in range(K):
        sched.add(A[j % 3], acts[j])
    assert len(sched) == K and A[0].step_index == K // 3
    sched.run(0, 11)
    sched.run()
    with pytest.raises(g.G2048Error):
        sched.run(0, 5)
    for j in range(K):
        B[j % 3].step(acts[j])
    for a, b in zip(A, B):
        assert torch.equal(a.boards, b.boards) and torch.equal(a.rewards, b.rewards) and torch.equal(a._dones, b._dones)
    assert torch.equal(A[1].legal_mask, B[1].legal_mask)


@pytest.mark.parametrize("policy,n,T,kw", [
    ("uniform", 5000, 30, {}),
    ("legal", 70001, 120, dict(illegal_move_reward=-1.0)),
    ("legal", 1 << 18, 40, dict(env_id_base=2**32 - 1000)),      # BASELINE config 4's size; ids cross 2^32
])
def test_step_with_in_kernel_policy_is_sample_actions_plus_step(policy, n, T, kw):
    """step(policy=...) — action drawn and played in ONE launch — == sample_actions() followed by step(), and the
    oracle's rollout under those actions; the legal policy never plays an illegal move while one is legal."""
    import torch
    import gym_2048_b200 as g
    outs = ("legal_mask", "illegal") if policy == "legal" else ()
    mk = lambda: g.BatchedGame2048(n, seed=31, outputs=outs, **kw)         # noqa: E731
    a, b = mk(), mk()
    a.reset(), b.reset()
    ref = oracle.OracleBatch(n, seed=31, threads=4, env_id_base=kw.get("env_id_base", 0),
                             illegal_move_reward=kw.get("illegal_move_reward", 0.0))
    ref.reset()
    n_illegal = 0
    for t in range(T):
        ra = a.step(policy=policy)
        act = b.sample_actions(legal=(policy == "legal"))
        rb = b.step(act)
        assert torch.equal(ra.actions, act), t
        assert torch.equal(ra.boards, rb.boards) and torch.equal(ra.rewards, rb.rewards) and torch.equal(ra.dones, rb.dones)
        o = ref.step(act.cpu().numpy())
        assert np.array_equal(ra.boards.cpu().numpy(), ref.boards)
        if policy == "legal":
            assert torch.equal(ra.legal_mask, rb.legal_mask)
            assert np.array_equal(ra.legal_mask.cpu().numpy(), o["legal_mask"])
            n_illegal += int(ra.illegal.sum())
    assert policy != "legal" or n_illegal == 0


def test_step_schedule_timed_run_records_events_between_regions():
    """StepSchedule.run(events=..., every=K): same boards as the untimed run, R + 1 events recorded in order."""
    import torch
    import gym_2048_b200 as g
    n, K, R = 20000, 5, 4
    a, b = (g.BatchedGame2048(n, seed=8, outputs=()) for _ in range(2))
    a.reset(), b.reset()
    acts = torch.randint(0, 4, (K * R, n), device="cuda", dtype=torch.uint8, generator=torch.Generator(device="cuda").manual_seed(3))
    sa, sb = g.StepSchedule(), g.StepSchedule()
    for j in range(K * R):
        sa.add(a, acts[j])
        sb.add(b, acts[j])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    sa.run(events=ev, every=K)
    sb.run()
    torch.cuda.synchronize()
    assert torch.equal(a.boards, b.boards)
    ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(R)]
    assert all(m > 0 for m in ms)
