"""Backend-independent parity checks (golden vectors + oracle cross-checks).

Each function takes an `ops` object (see backends.py: .Batch, .move, .status, .philox)
and asserts bit-exact agreement.  Used by the CPU suite (oracle, host simulation of the
device code) and by the GPU suite (the product through the C ABI)."""
import numpy as np

from conftest import DONE_KEYS, STEP_KEYS, load_golden, rollout_cfg
from oracle import oracle


def check_shift_table_all_directions(ops):
    z = load_golden("shift_table.npz")
    rows, want, score = z["rows"], z["out"], z["score"]
    n = len(rows)
    for d in range(4):
        vertical, rev = d in (0, 2), d in (1, 2)
        src = rows[:, ::-1] if rev else rows
        dst = want[:, ::-1] if rev else want
        b = np.zeros((n, 4, 4), np.uint8)
        w = np.zeros((n, 4, 4), np.uint8)
        line = np.arange(n) % 4
        if vertical:
            b[np.arange(n), :, line] = src
            w[np.arange(n), :, line] = dst
        else:
            b[np.arange(n), line, :] = src
            w[np.arange(n), line, :] = dst
        out, s, ch = ops.move(b.reshape(n, 16), np.full(n, d, np.uint8))
        assert np.array_equal(out, w.reshape(n, 16)), d
        assert np.array_equal(s, score), d
        assert np.array_equal(ch != 0, (rows != want).any(axis=1)), d


def check_csv_transitions(ops):
    z = load_golden("csv_transitions.npz")
    n = len(z["boards"])
    for max_exp, key in ((0, "terminated_max_none"), (11, "terminated_max_2048")):
        b = ops.Batch(n, auto_reset=False, max_tile_exp=max_exp)
        b.boards[:] = z["boards"]
        out = b.step(z["actions"], forced_draws=z["words"])
        assert np.array_equal(b.boards, z["next_boards"])
        assert np.array_equal(out["rewards"], z["rewards"])
        assert np.array_equal(out["dones"], z[key])
        assert np.array_equal(out["highest_exp"], z["highest_exp"])
        assert not out["illegal"].any()


def check_special_boards(ops):
    z = load_golden("special.npz")
    boards = z["boards"]
    n = len(boards)
    st = ops.status(boards)
    assert np.array_equal(st["legal_mask"], z["legal_mask"])
    assert np.array_equal(st["n_empty"], z["n_empty"])
    assert np.array_equal(st["highest_exp"], z["board_highest_exp"])
    assert np.array_equal(st["is_end"], z["is_end"][:, 0])
    assert np.array_equal(ops.status(boards, max_tile_exp=11)["is_end"], z["is_end"][:, 1])
    for d in range(4):
        out, s, ch = ops.move(boards, np.full(n, d, np.uint8))
        assert np.array_equal(out, z["move_boards"][:, d])
        assert np.array_equal(ch, z["move_changed"][:, d])
        assert np.array_equal(s[ch != 0], z["move_scores"][:, d][ch != 0])
        b = ops.Batch(n, auto_reset=False, illegal_move_reward=-1.0)
        b.boards[:] = boards
        o = b.step(np.full(n, d, np.uint8), forced_draws=z["words"][:, d])
        v = z["valid"][:, d] != 0
        for k, g in (("rewards", "rewards"), ("dones", "dones"), ("illegal", "illegal"),
                     ("highest_exp", "highest_exp")):
            assert np.array_equal(o[k][v], z[g][:, d][v]), (d, k)
        assert np.array_equal(b.boards[v], z["out_boards"][:, d][v])


def check_rollouts(ops, rollouts, max_steps=None):
    for name, rec in rollouts.items():
        cfg = rollout_cfg(rec)
        T = cfg.pop("T")
        if max_steps:
            T = min(T, max_steps)
        b = ops.Batch(**cfg)
        assert np.array_equal(b.reset(), rec["init_boards"]), name
        ep_ret = np.zeros(cfg["n"], np.float32)        # SB3 Monitor's running sum, from the REFERENCE's own rewards
        for t in range(T):
            out = b.step(rec["actions"][t])
            out["ep_score"], out["ep_len"] = b.ep_score, b.ep_len
            for k in STEP_KEYS:
                assert np.array_equal(out[k], rec[k][t]), (name, t, k)
            d = rec["dones"][t] != 0
            for k in DONE_KEYS:
                assert np.array_equal(out[k][d], rec[k][t][d]), (name, t, k)
            ep_ret += rec["rewards"][t]
            assert np.array_equal(out["final_return"][d], ep_ret[d]), (name, t, "final_return")
            if cfg["auto_reset"]:
                ep_ret[d] = 0
            assert np.array_equal(b.ep_return, ep_ret), (name, t, "ep_return")


def check_against_oracle(ops, n=4096, steps=64, seed=7, policy="random", env_id_base=0, max_tile_exp=0,
                         auto_reset=True, illegal_move_reward=0.0, threads=4):
    """Seeded rollout of `ops` next to the C oracle, every output compared every step."""
    kw = dict(seed=seed, env_id_base=env_id_base, max_tile_exp=max_tile_exp, auto_reset=auto_reset,
              illegal_move_reward=illegal_move_reward)
    a, b = oracle.OracleBatch(n, threads=threads, **kw), ops.Batch(n, **kw)
    assert np.array_equal(a.reset(), b.reset())
    rng = np.random.default_rng(seed + 1)
    mask = oracle.status(a.boards)["legal_mask"]
    for t in range(steps):
        if policy == "legal":
            # uniform among legal moves: rotate a random start until a legal bit is hit
            act = rng.integers(0, 4, n).astype(np.uint8)
            for _ in range(3):
                bad = ((mask >> act) & 1) == 0
                act[bad] = (act[bad] + 1) & 3
        else:
            act = rng.integers(0, 4, n).astype(np.uint8)
        oa, ob = a.step(act), b.step(act)
        oa["ep_score"], oa["ep_len"], ob["ep_score"], ob["ep_len"] = a.ep_score, a.ep_len, b.ep_score, b.ep_len
        oa["ep_return"], ob["ep_return"] = a.ep_return, b.ep_return
        for k in STEP_KEYS + ("ep_return",):
            assert np.array_equal(oa[k], ob[k]), (t, k)
        d = oa["dones"] != 0
        for k in DONE_KEYS + ("final_return",):
            assert np.array_equal(oa[k][d], ob[k][d]), (t, k)
        mask = oa["legal_mask"]
    return a


def check_status_random_boards(ops, n=20000, seed=3):
    rng = np.random.default_rng(seed)
    hi = rng.integers(1, 18, size=(n, 1))
    e = rng.integers(1, 18, size=(n, 16)) % (hi + 1)
    e[rng.random((n, 16)) < rng.choice([0.0, 0.1, 0.5], size=(n, 1))] = 0
    boards = e.astype(np.uint8)
    for mt in (0, 3, 11):
        sa, sb = oracle.status(boards, mt), ops.status(boards, mt)
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), (mt, k)
    for d in range(4):
        ra, rb = oracle.move(boards, np.full(n, d, np.uint8)), ops.move(boards, np.full(n, d, np.uint8))
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y), d


def check_add_tile(ops, n=20000, seed=5):
    rng = np.random.default_rng(seed)
    e = rng.integers(1, 12, size=(n, 16))
    e[rng.random((n, 16)) < rng.choice([0.0, 0.07, 0.3, 0.9, 1.0], size=(n, 1))] = 0
    boards = e.astype(np.uint8)
    for idx in (0, 1, (1 << 32) + 7):
        a = oracle.add_tile(boards, 1000, seed, idx)
        b = ops.add_tile(boards, 1000, seed, idx)
        assert np.array_equal(a, b)
        changed = (a != boards).sum(axis=1)
        assert np.array_equal(changed, ((boards == 0).sum(axis=1) > 0).astype(int))


def check_philox(ops):
    from test_draws import KATS
    for ctr, key, want in KATS:
        assert tuple(int(x) for x in ops.philox([ctr], key[0], key[1])[0]) == want
    rng = np.random.default_rng(11)
    ctr = rng.integers(0, 2**32, size=(2048, 4), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(ops.philox(ctr, 5, 9), oracle.philox(ctr, 5, 9))
    from test_draws import KATS_2X32
    for ctr2, key, want in KATS_2X32:
        assert tuple(int(x) for x in ops.philox2x32([ctr2], key)[0]) == want
    ctr2 = rng.integers(0, 2**32, size=(2048, 2), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(ops.philox2x32(ctr2, 0xDEADBEEF), oracle.philox2x32(ctr2, 0xDEADBEEF))


def check_draw_words(ops):
    """The draw words of include/g2048.h "Draw stream" for every tag, across a 2^32 env-id
    boundary and with 64-bit seeds and indices, against the numpy statement of the spec."""
    from oracle import draws
    for seed, idx, base in ((0, 0, 0), (42, 5, 1000), (2**64 - 1, 2**64 - 1, 2**32 - 700),
                            (0x123456789ABCDEF, 2**40 + 7, 2**63 + 2**32 - 3)):
        for tag in (0, 1, 2):
            n = 1500
            env = np.arange(n, dtype=np.uint64) + np.uint64(base)
            want = draws.draw_words(seed, env, idx, tag)
            assert np.array_equal(ops.draw_words(n, base, seed, idx, tag), want), (seed, idx, base, tag)
