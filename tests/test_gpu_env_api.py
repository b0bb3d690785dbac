"""The single-env class, observations, VecEnv adapter and host-buffer handle on the GPU.

TestBoard / TestStep restate the reference's env/envs/test_game2048_env.py known answers
(file:line in comments) against gym_2048_b200.Game2048Env, so a user of the reference class
finds the same behaviour."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

MOVE_BOARD = [[0, 2, 0, 4], [2, 2, 8, 0], [2, 2, 2, 8], [2, 2, 4, 4]]
DEAD = [[2, 4, 8, 16], [4, 8, 16, 2], [8, 16, 2, 4], [16, 2, 4, 8]]


@pytest.fixture(scope="module")
def g():
    import gym_2048_b200
    return gym_2048_b200


class TestBoard:
    def test_shift(self, g):                                     # test_game2048_env.py:10-34
        b = g.Game2048Env()
        assert b.shift([0, 0, 0, 0]) == ([0, 0, 0, 0], 0)
        assert b.shift([0, 2, 0, 0]) == ([2, 0, 0, 0], 0)
        assert b.shift([0, 2, 0, 4]) == ([2, 4, 0, 0], 0)
        assert b.shift([2, 4, 8, 16]) == ([2, 4, 8, 16], 0)
        assert b.shift([2, 2, 8, 0]) == ([4, 8, 0, 0], 4)
        assert b.shift([4, 2, 2, 4]) == ([4, 4, 4, 0], 4)
        assert b.shift([2, 2, 2, 8]) == ([4, 2, 8, 0], 4)
        assert b.shift([2, 8, 4, 4]) == ([2, 8, 8, 0], 8)
        assert b.shift([2, 2, 4, 4]) == ([4, 8, 0, 0], 12)
        assert b.shift([2, 4, 4, 4]) == ([2, 8, 4, 0], 8)
        assert b.shift([4, 4, 4, 4]) == ([8, 8, 0, 0], 16)
        assert b.shift([0, 2, 2, 8]) == ([4, 8, 0, 0], 4)

    def test_move(self, g):                                      # :36-98
        b = g.Game2048Env()
        want = {0: ([[4, 4, 8, 4], [2, 4, 2, 8], [0, 0, 4, 4], [0, 0, 0, 0]], 12),
                1: ([[0, 0, 2, 4], [0, 0, 4, 8], [0, 2, 4, 8], [0, 0, 4, 8]], 20),
                2: ([[0, 0, 0, 0], [0, 0, 8, 4], [2, 4, 2, 8], [4, 4, 4, 4]], 12),
                3: ([[2, 4, 0, 0], [4, 8, 0, 0], [4, 2, 8, 0], [4, 8, 0, 0]], 20)}
        for d in (0, 1, 2, 3):
            b.set_board(np.array(MOVE_BOARD))
            assert b.move(d) == want[d][1]
            assert np.array_equal(b.get_board(), np.array(want[d][0]))
        with pytest.raises(g.IllegalMove):                       # :89-90
            b.move(3)
        assert b.move(2) == 8                                    # :93-98
        assert np.array_equal(b.get_board(), np.array([[0, 4, 0, 0], [2, 8, 0, 0], [4, 2, 0, 0], [8, 8, 8, 0]]))

    def test_trial_move_leaves_board(self, g):
        b = g.Game2048Env()
        b.set_board(np.array(MOVE_BOARD))
        assert b.move(0, trial=True) == 12
        assert np.array_equal(b.get_board(), np.array(MOVE_BOARD))

    def test_set_board_aliases_callers_array(self, g):
        b = g.Game2048Env()
        arr = np.array(MOVE_BOARD)
        b.set_board(arr)
        b.move(0)
        assert b.get_board() is arr and arr[0, 0] == 4

    def test_highest(self, g):                                   # :100-107
        b = g.Game2048Env()
        b.set_board(np.array([[0, 2, 0, 4], [2, 2, 8, 0], [2, 2, 2048, 8], [2, 2, 4, 4]]))
        assert b.highest() == 2048

    def test_isend(self, g):                                     # :109-151
        b = g.Game2048Env()
        b.set_board(np.array([[2] * 4] * 4))
        assert b.isend() == False                                # noqa: E712
        b.set_board(np.array(DEAD))
        assert b.isend() == True                                 # noqa: E712
        hole = np.array(DEAD)
        hole[3, 3] = 0
        b.set_board(hole)
        assert b.isend() == False                                # noqa: E712
        b.set_max_tile(2048)
        lone = np.zeros((4, 4), int)
        lone[0, 0] = 2048
        b.set_board(lone)
        assert b.isend() == True                                 # noqa: E712
        lone[0, 0] = 1024
        b.set_board(lone)
        assert b.isend() == False                                # noqa: E712


class TestStep:
    def test_step_returns_correct_shapes(self, g):               # :154-163
        b = g.Game2048Env()
        b.reset(seed=0)
        obs, reward, terminated, truncated, info = b.step(0)
        assert obs.shape == (16, 4, 4)
        assert isinstance(reward, float) and isinstance(terminated, bool) and isinstance(truncated, bool)
        assert 'illegal_move' in info and 'highest' in info

    def test_step_reward_and_score(self, g):                     # :165-192
        b = g.Game2048Env()
        b.reset(seed=0)
        b.set_board(np.array([[0] * 4, [0] * 4, [2, 0, 0, 0], [2, 0, 0, 0]]))
        assert b.step(0)[1] == 4.0
        b.set_board(np.array([[0] * 4, [0] * 4, [4, 0, 0, 0], [4, 0, 0, 0]]))
        b.step(0)
        assert b.score == 12.0

    def test_step_illegal_move(self, g):                         # :194-217
        b = g.Game2048Env()
        b.reset(seed=0)
        b.set_board(np.array(DEAD))
        obs, reward, terminated, truncated, info = b.step(0)
        assert terminated == True and info['illegal_move'] == True and reward == 0.0   # noqa: E712
        assert np.array_equal(b.get_board(), np.array(DEAD))
        b = g.Game2048Env()
        b.set_illegal_move_reward(-1.0)
        b.reset(seed=0)
        b.set_board(np.array(DEAD))
        assert b.step(0)[1] == -1.0

    def test_step_observation_is_valid_one_hot(self, g):         # :219-231
        b = g.Game2048Env()
        b.reset(seed=0)
        b.set_board(np.array([[2, 0, 0, 0], [0] * 4, [0] * 4, [0, 0, 4, 0]]))
        obs = b.step(1)[0]
        assert obs.shape == (16, 4, 4) and obs.sum(axis=0).max() <= 1
        assert set(obs.flatten().tolist()) == {0, 1}
        assert obs.dtype == np.int64

    def test_reset_is_seed_deterministic_and_matches_oracle(self, g):
        a, b = g.Game2048Env(), g.Game2048Env()
        oa, _ = a.reset(seed=456)
        ob, _ = b.reset(seed=456)
        assert np.array_equal(oa, ob) and (a.get_board() != 0).sum() == 2 and a.score == 0
        o = oracle.OracleBatch(1, seed=456, auto_reset=False)
        assert np.array_equal(oracle.exp_to_values(o.reset()[0]).reshape(4, 4), a.get_board())
        rng = np.random.default_rng(0)
        for _ in range(60):                                      # same game as the oracle, step by step
            act = int(rng.integers(4))
            out = o.step([act])
            obs, reward, terminated, _, info = a.step(act)
            assert np.array_equal(oracle.exp_to_values(o.boards[0]).reshape(4, 4), a.get_board())
            assert reward == out["rewards"][0] and terminated == bool(out["dones"][0])
            assert info["illegal_move"] == bool(out["illegal"][0])
            assert info["highest"] == (1 << int(out["highest_exp"][0]))
            if terminated:
                break

    def test_stack_function(self, g):                            # reference stack :17-32
        z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "special.npz"))
        for i in (0, 3, 7, 8, 9):
            vals = oracle.exp_to_values(z["boards"][i]).reshape(4, 4)
            assert np.array_equal(g.stack(vals), z["obs"][i].astype(int))


def test_observe_all_dtypes_match_reference_stack(g):
    import torch
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "special.npz"))
    boards = z["boards"]
    game = g.BatchedGame2048(len(boards), outputs=())
    game.set_boards(torch.from_numpy(boards))
    for dt in (torch.uint8, torch.float32, torch.int64, torch.bfloat16):
        obs = game.observe(dt)
        assert obs.dtype == dt and tuple(obs.shape) == (len(boards), 16, 4, 4)
        assert np.array_equal(obs.float().cpu().numpy().astype(np.uint8), z["obs"])
    vals = game.board_values().cpu().numpy().reshape(-1, 16)
    assert np.array_equal(vals, oracle.exp_to_values(boards))
    game2 = g.BatchedGame2048(len(boards), outputs=())
    game2.set_board_values(torch.from_numpy(vals))
    assert torch.equal(game2.boards, game.boards)
    with pytest.raises(ValueError):
        game2.set_board_values(torch.full((len(boards), 16), 3))


def test_value_exponent_converters_rows_and_ragged(g):
    """g2048_values_from_exp / g2048_exp_from_values: the row-per-thread path (whole, aligned rows) and the
    per-cell path (ragged counts, unaligned pointers) give the same cells; bad values are counted."""
    import ctypes as C
    import torch
    from gym_2048_b200._lib import check, lib
    L = lib()
    rng = np.random.default_rng(5)
    exps = rng.integers(0, 19, 4096 + 64).astype(np.uint8)
    want = np.where(exps > 0, np.int64(1) << exps.astype(np.int64), 0)
    d_exp = torch.from_numpy(exps).cuda()
    for off, cnt in ((0, 4096), (0, 4095), (1, 4093), (4, 4000), (3, 7), (0, 1)):
        vals = torch.full((4200,), -1, dtype=torch.int64, device="cuda")
        check(L.g2048_values_from_exp(C.c_void_p(d_exp.data_ptr() + off), C.c_void_p(vals.data_ptr()), cnt, None))
        assert np.array_equal(vals[:cnt].cpu().numpy(), want[off:off + cnt]) and int(vals[cnt]) == -1
        back = torch.full((4200,), 255, dtype=torch.uint8, device="cuda")
        bad = torch.zeros(1, dtype=torch.int32, device="cuda")
        check(L.g2048_exp_from_values(C.c_void_p(vals.data_ptr()), C.c_void_p(back.data_ptr() + off), cnt,
                                      C.c_void_p(bad.data_ptr()), None))
        assert np.array_equal(back[off:off + cnt].cpu().numpy(), exps[off:off + cnt]) and int(back[off + cnt]) == 255
        assert int(bad) == 0
    vals = torch.tensor([2, 3, 0, 6, 4, 1, 1 << 31, 1 << 32], dtype=torch.int64, device="cuda")
    out = torch.empty(8, dtype=torch.uint8, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    check(L.g2048_exp_from_values(C.c_void_p(vals.data_ptr()), C.c_void_p(out.data_ptr()), 8, C.c_void_p(bad.data_ptr()), None))
    assert out.cpu().tolist() == [1, 0, 0, 0, 2, 0, 31, 0] and int(bad) == 4


@pytest.mark.parametrize("penalty", [0.0, -1.0, -2.5])
def test_vec_env_sb3_semantics(g, penalty):
    """SB3 DummyVecEnv + Monitor semantics vs the oracle; `episode.r` is Monitor's sum of the rewards the agent
    saw — illegal-move penalty included (a negative return is representable) — not the game score."""
    n = 512
    venv = g.Game2048VecEnv(n, seed=7, obs_dtype=__import__("torch").uint8, illegal_move_reward=penalty)
    o = oracle.OracleBatch(n, seed=7, illegal_move_reward=penalty)
    obs = venv.reset()
    assert obs.shape == (n, 16, 4, 4) and np.array_equal(obs, oracle.encode_obs_u8(o.reset()))
    assert venv.observation_space.shape == (16, 4, 4) and venv.action_space.n == 4
    rng = np.random.default_rng(3)
    n_done, n_neg = 0, 0
    ep_ret = np.zeros(n, np.float64)
    for t in range(60):
        act = rng.integers(0, 4, n)
        venv.step_async(act)
        obs, rew, dones, infos = venv.step_wait()
        out = o.step(act.astype(np.uint8))
        assert rew.dtype == np.float32 and dones.dtype == np.bool_ and len(infos) == n
        assert np.array_equal(obs, oracle.encode_obs_u8(o.boards))
        assert np.array_equal(rew, out["rewards"]) and np.array_equal(dones, out["dones"] != 0)
        ep_ret += out["rewards"]
        for i in np.flatnonzero(dones):
            info = infos[i]
            assert info["TimeLimit.truncated"] is False
            assert np.array_equal(info["terminal_observation"], oracle.encode_obs_u8(out["terminal_boards"][i])[0])
            assert info["episode"]["r"] == round(float(ep_ret[i]), 6) and info["episode"]["l"] == int(out["final_len"][i])
            assert info["score"] == int(out["final_score"][i])
            assert info["highest"] == (1 << int(out["highest_exp"][i])) and info["illegal_move"] == bool(out["illegal"][i])
            n_done += 1
            n_neg += info["episode"]["r"] < 0
        ep_ret[dones] = 0.0
        masks = venv.action_masks()
        assert np.array_equal(masks, ((out["legal_mask"][:, None] >> np.arange(4)) & 1).astype(bool))
    assert n_done > 50
    assert penalty == 0.0 or n_neg > 0


@pytest.mark.parametrize("wire", ["packed", "plain"])    # packed (default): the boards cross PCIe 4 bits per cell
@pytest.mark.parametrize("extras", [True, False])       # False: the O_EPRUN kernel bench.py's e2e leg runs
def test_host_stepped_env_matches_oracle(g, extras, wire):
    n = 20000
    h = g.HostSteppedEnv(n, seed=9, n_chunks=3, extras=extras, wire=wire)
    o = oracle.OracleBatch(n, seed=9, threads=4)
    assert np.array_equal(h.reset().numpy(), o.reset())
    rng = np.random.default_rng(1)
    for t in range(30):
        act = rng.integers(0, 4, n).astype(np.uint8)
        b = h.step(act)
        out = o.step(act)
        assert np.array_equal(b.boards.numpy(), o.boards)
        assert np.array_equal(b.rewards.numpy(), out["rewards"]) and np.array_equal(b.dones.numpy(), out["dones"])
        if extras:
            assert np.array_equal(b.illegal.numpy(), out["illegal"])
            assert np.array_equal(b.highest_exp.numpy(), out["highest_exp"])
            assert np.array_equal(b.legal_mask.numpy(), out["legal_mask"])
    assert h.step_index == 30
    h.close()


def test_host_stepped_env_nibble_boards(g):
    """board_format='nibble': the same step, the boards coming back 4 bits per cell (13 instead of 21 bytes per board
    over PCIe); boards that do not fit are counted and the full boards stay available."""
    import torch
    n = 30000
    h = g.HostSteppedEnv(n, seed=9, n_chunks=3, board_format="nibble")
    o = oracle.OracleBatch(n, seed=9, threads=4)
    assert np.array_equal(h.reset().numpy(), o.reset())
    rng = np.random.default_rng(1)
    for t in range(25):
        act = rng.integers(0, 4, n).astype(np.uint8)
        b = h.step(act)
        out = o.step(act)
        assert b.boards.shape == (n, 8)
        assert np.array_equal(b.unpacked_boards().numpy(), o.boards)
        assert np.array_equal(b.rewards.numpy(), out["rewards"]) and np.array_equal(b.dones.numpy(), out["dones"])
        assert b.nibble_overflow.value == 0
    assert np.array_equal(h.full_boards().numpy(), o.boards)
    # boards with tiles >= 65,536 do not fit 4 bits: counted, and the 16-byte boards are still there
    big = o.boards.copy()
    big[:7, 0] = 16
    big[:7, 1] = 17
    big[:7, 2:] = 0
    check = g._lib.check
    check(h.lib.g2048_env_set_boards_host(h._h, big.ctypes.data))
    o.boards[:] = big
    act = np.full(n, 2, np.uint8)                      # Down: the two big tiles stay on the board
    b = h.step(act)
    o.step(act)
    assert b.nibble_overflow.value == 7
    assert np.array_equal(h.full_boards().numpy(), o.boards)
    assert np.array_equal(b.unpacked_boards().numpy()[7:], o.boards[7:])
    h.close()


@pytest.mark.parametrize("n,n_chunks,threads", [(1, 0, 1), (255, 0, 3), (70001, 0, 0), (300000, 7, 5), (1 << 20, 0, 0)])
def test_host_stepped_env_packed_wire_equals_plain_wire(g, n, n_chunks, threads):
    """G2048_BOARDS_BYTES_PACKED_WIRE: [n,16] exponent bytes exactly as the plain format delivers them — odd sizes,
    slice counts and thread counts — and a step with tiles >= 65,536 (which do not fit 4 bits) still comes back whole."""
    a = g.HostSteppedEnv(n, seed=3, n_chunks=n_chunks, wire="packed", unpack_threads=threads)
    b = g.HostSteppedEnv(n, seed=3, n_chunks=2, wire="plain")
    assert a.wire == "packed" and b.wire == "plain"
    assert np.array_equal(a.reset().numpy(), b.reset().numpy())
    rng = np.random.default_rng(5)
    for t in range(12):
        act = rng.integers(0, 4, n).astype(np.uint8)
        ra, rb = a.step(act), b.step(act)
        assert np.array_equal(ra.boards.numpy(), rb.boards.numpy()), t
        assert np.array_equal(ra.rewards.numpy(), rb.rewards.numpy()) and np.array_equal(ra.dones.numpy(), rb.dones.numpy())
        assert ra.nibble_overflow.value == 0
    big = rb.boards.numpy().copy()
    k = min(n, 5)
    big[:k, 0], big[:k, 1], big[:k, 2:] = 16, 17, 0
    for h in (a, b):
        g._lib.check(h.lib.g2048_env_set_boards_host(h._h, big.ctypes.data))
    act = np.full(n, 2, np.uint8)                      # Down: the big tiles stay on the board
    ra, rb = a.step(act), b.step(act)
    assert ra.nibble_overflow.value == k
    assert np.array_equal(ra.boards.numpy(), rb.boards.numpy()) and int(ra.boards.numpy().max()) == 17
    a.close(), b.close()


def test_checkpoint_resume_is_exact(g):
    import torch
    n = 4096
    a = g.BatchedGame2048(n, seed=5)
    a.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.uint8) for _ in range(40)]
    for t in range(20):
        a.step(acts[t])
    sd = a.state_dict()
    b = g.BatchedGame2048(n, seed=999)
    b.load_state_dict(sd)
    for t in range(20, 40):
        ra, rb = a.step(acts[t]), b.step(acts[t])
        assert torch.equal(ra.boards, rb.boards) and torch.equal(ra.rewards, rb.rewards)
        assert torch.equal(a.ep_score, b.ep_score)


def test_argument_validation(g):
    import ctypes as C
    import torch
    game = g.BatchedGame2048(64)
    game.reset()
    with pytest.raises(ValueError):
        game.step(torch.full((64,), 4))
    with pytest.raises(ValueError):
        game.step(torch.zeros(63, dtype=torch.uint8))
    L = g._lib.lib()
    assert L.g2048_step(None, None) == -1 and b"NULL" in L.g2048_last_error()
    a = g._lib.StepArgs()
    a.n = 4
    assert L.g2048_step(C.byref(a), None) == -1
    buf = torch.zeros(4 * 16 + 1, dtype=torch.uint8, device="cuda")
    assert L.g2048_reset(C.c_void_p(buf.data_ptr() + 1), None, 4, 0, 0, 0, None) == -2
    assert L.g2048_encode_obs(C.c_void_p(buf.data_ptr()), C.c_void_p(buf.data_ptr()), 99, 1, None) == -1


def test_bench_prints_one_contract_line_on_a_small_workload():
    """bench.py on a small batch (fast): ONE JSON line with the contract's keys, the step-list issue path for a
    shard the host cannot keep up with, and a state checksum that does not depend on how the run is timed."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(*extra):
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--envs", "65536", "--steps", "4", "--warmup", "3",
                            "--repeats", "3", "--spinup", "64", "--sets", "4", "--e2e-steps", "3", "--fused-steps", "4",
                            "--no-cpu-baseline", "--no-config4"] + list(extra), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1, r.stdout
        return json.loads(lines[0])
    d = run()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "repeats", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "e2e_compact", "gpu_launches", "clocks",
              "state_checksum", "timing"):
        assert k in d, k
    assert d["metric"] == "env_steps_per_sec" and d["scaling"] == "strong" and d["dtype"] == "u8" and d["value"] > 1e9
    assert d["config"]["global_envs"] == 65536 and d["gpu_launches"] == 12 and d["repeats"] == 3
    assert "g2048_step_list" in d["timing"]["issue"] and "chained launches" in d["timing"]["issue"]
    assert d["timing"]["issue_threads"] == 2 and d["plain_launches"]["value"] > 1e9 and d["long_region"]["launches"] == 1000
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic_kind"]
    assert abs(r["achieved"] - 38 * 65536 / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert d["e2e"]["d2h_bytes_per_step"] == 65536 * 13 and d["e2e_compact"]["d2h_bytes_per_step"] == 65536 * 13
    assert d["e2e_plain_wire"]["d2h_bytes_per_step"] == 65536 * 21
    assert d["e2e"]["checksum"] == d["e2e_compact"]["checksum"] == d["e2e_plain_wire"]["checksum"]     # same rewards through every host format
    assert d["e2e"]["boards_checksum"] == d["e2e_plain_wire"]["boards_checksum"]      # and the same boards, packed or plain on the wire
    # the Python-loop issue path (forced) steps the same boards: identical checksum
    d2 = run("--small-below", "0")
    assert "Python loop" in d2["timing"]["issue"] and d2["state_checksum"] == d["state_checksum"]
    # plain launches from one issuing thread (the first half of round 2): the same boards again
    d3 = run("--chain", "off", "--issue-threads", "1")
    assert "plain launches" in d3["timing"]["issue"] and d3["plain_launches"] is None
    assert d3["state_checksum"] == d["state_checksum"]
