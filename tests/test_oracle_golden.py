"""Pins the C oracle (oracle/g2048_oracle.c) to the reference.

Three sources, all committed under tests/golden/ by make_golden.py (which ran the
unmodified reference env): the reference's own unit-test known answers
(env/envs/test_game2048_env.py, restated below with file:line), its
data/test_data.csv transitions, and injected-draw rollouts / special boards.
CPU only.
"""
import numpy as np
import pytest

from conftest import DONE_KEYS, STEP_KEYS, load_golden, rollout_cfg
from oracle import oracle


def E(rows):
    return oracle.values_to_exp(np.array(rows, dtype=np.int64).reshape(-1))


# -- reference unit-test known answers ---------------------------------------------------
SHIFT_KAT = [  # test_game2048_env.py:13-34
    ([0, 0, 0, 0], [0, 0, 0, 0], 0), ([0, 2, 0, 0], [2, 0, 0, 0], 0), ([0, 2, 0, 4], [2, 4, 0, 0], 0),
    ([2, 4, 8, 16], [2, 4, 8, 16], 0), ([2, 2, 8, 0], [4, 8, 0, 0], 4), ([4, 2, 2, 4], [4, 4, 4, 0], 4),
    ([2, 2, 2, 8], [4, 2, 8, 0], 4), ([2, 8, 4, 4], [2, 8, 8, 0], 8), ([2, 2, 4, 4], [4, 8, 0, 0], 12),
    ([2, 4, 4, 4], [2, 8, 4, 0], 8), ([4, 4, 4, 4], [8, 8, 0, 0], 16), ([0, 2, 2, 8], [4, 8, 0, 0], 4),
]
MOVE_BOARD = [[0, 2, 0, 4], [2, 2, 8, 0], [2, 2, 2, 8], [2, 2, 4, 4]]  # :40-44
MOVE_KAT = {  # :45-86
    0: ([[4, 4, 8, 4], [2, 4, 2, 8], [0, 0, 4, 4], [0, 0, 0, 0]], 12),
    1: ([[0, 0, 2, 4], [0, 0, 4, 8], [0, 2, 4, 8], [0, 0, 4, 8]], 20),
    2: ([[0, 0, 0, 0], [0, 0, 8, 4], [2, 4, 2, 8], [4, 4, 4, 4]], 12),
    3: ([[2, 4, 0, 0], [4, 8, 0, 0], [4, 2, 8, 0], [4, 8, 0, 0]], 20),
}
DEAD = [[2, 4, 8, 16], [4, 8, 16, 2], [8, 16, 2, 4], [16, 2, 4, 8]]  # :121-125


def test_shift_known_answers():
    for row, want, score in SHIFT_KAT:
        got, s = oracle.shift(E(row))
        assert (list(oracle.exp_to_values(got)), s) == (want, score)


def test_move_known_answers():
    for d, (want, score) in MOVE_KAT.items():
        out, s, ch = oracle.move(E(MOVE_BOARD)[None], [d])
        assert np.array_equal(out[0], E(want)) and s[0] == score and ch[0] == 1
    left = E(MOVE_KAT[3][0])
    out, s, ch = oracle.move(left[None], [3])            # :89-90 repeat move is illegal
    assert ch[0] == 0 and np.array_equal(out[0], left)
    out, s, ch = oracle.move(left[None], [2])            # :93-98 follow-on move
    assert s[0] == 8 and np.array_equal(out[0], E([[0, 4, 0, 0], [2, 8, 0, 0], [4, 2, 0, 0], [8, 8, 8, 0]]))


def test_highest_and_isend_known_answers():
    st = oracle.status(E([[0, 2, 0, 4], [2, 2, 8, 0], [2, 2, 2048, 8], [2, 2, 4, 4]])[None])
    assert st["highest_exp"][0] == 11                                      # :100-107
    assert oracle.status(E([[2] * 4] * 4)[None])["is_end"][0] == 0         # :113-118
    assert oracle.status(E(DEAD)[None])["is_end"][0] == 1                  # :121-126
    hole = [r[:] for r in DEAD]
    hole[3][3] = 0
    assert oracle.status(E(hole)[None])["is_end"][0] == 0                  # :129-134
    lone = lambda v: E([[v, 0, 0, 0]] + [[0] * 4] * 3)[None]               # noqa: E731
    assert oracle.status(lone(2048), max_tile_exp=11)["is_end"][0] == 1    # :137-143
    assert oracle.status(lone(1024), max_tile_exp=11)["is_end"][0] == 0    # :146-151


def _one_step(board_rows, action, **kw):
    b = oracle.OracleBatch(1, auto_reset=False, **kw)
    b.boards[0] = E(board_rows)
    return b, b.step([action])


def test_step_known_answers():
    col = [[0] * 4, [0] * 4, [2, 0, 0, 0], [2, 0, 0, 0]]
    b, out = _one_step(col, 0)                                             # :165-175
    assert out["rewards"][0] == 4.0 and out["dones"][0] == 0 and out["illegal"][0] == 0
    b.boards[0] = E([[0] * 4, [0] * 4, [4, 0, 0, 0], [4, 0, 0, 0]])
    b.step([0])
    assert b.ep_score[0] == 12                                             # :177-192
    b, out = _one_step(DEAD, 0)                                            # :194-205
    assert out["dones"][0] == 1 and out["illegal"][0] == 1 and out["rewards"][0] == 0.0
    assert np.array_equal(b.boards[0], E(DEAD))
    b, out = _one_step(DEAD, 0, illegal_move_reward=-1.0)                  # :207-217
    assert out["rewards"][0] == -1.0
    obs = oracle.encode_obs_u8(E([[2, 0, 0, 0], [0] * 4, [0] * 4, [0, 0, 4, 0]])[None])[0]
    assert obs.shape == (16, 4, 4) and obs.sum(axis=0).max() <= 1          # :219-231
    assert set(obs.flatten().tolist()) == {0, 1}


# -- generated from the unmodified reference --------------------------------------------
def test_shift_table_exhaustive():
    z = load_golden("shift_table.npz")
    lib = oracle.lib()
    rows, want, score = z["rows"], z["out"], z["score"]
    import ctypes as C
    out = (C.c_uint8 * 4)()
    for i in range(0, len(rows)):
        r = (C.c_uint8 * 4)(*rows[i].tolist())
        s = lib.g2048_oracle_shift(r, out)
        assert s == score[i] and list(out) == want[i].tolist(), rows[i]


def test_shift_table_via_move_all_directions():
    """Each table row embedded as a line of a board, all 4 directions (orientation check)."""
    z = load_golden("shift_table.npz")
    rows, want, score = z["rows"], z["out"], z["score"]
    n = len(rows)
    for d in range(4):
        vertical = d in (0, 2)
        rev = d in (1, 2)
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
        out, s, ch = oracle.move(b.reshape(n, 16), np.full(n, d, np.uint8))
        assert np.array_equal(out, w.reshape(n, 16))
        assert np.array_equal(s, score)
        assert np.array_equal(ch != 0, (rows != want).any(axis=1))


@pytest.mark.parametrize("max_exp,key", [(0, "terminated_max_none"), (11, "terminated_max_2048")])
def test_csv_transitions(max_exp, key):
    z = load_golden("csv_transitions.npz")
    n = len(z["boards"])
    b = oracle.OracleBatch(n, auto_reset=False, max_tile_exp=max_exp)
    b.boards[:] = z["boards"]
    out = b.step(z["actions"], forced_draws=z["words"])
    assert np.array_equal(b.boards, z["next_boards"])
    assert np.array_equal(out["rewards"], z["rewards"])
    assert np.array_equal(out["dones"], z[key])
    assert np.array_equal(out["highest_exp"], z["highest_exp"])
    assert not out["illegal"].any()
    if max_exp == 11:
        assert np.array_equal(out["dones"], z["done_csv"])   # the CSV was recorded with max_tile=2048
    # the file is one chained game: next board of row i is the board of row i+1
    assert np.array_equal(z["next_boards"][:-1], z["boards"][1:])


def test_special_boards():
    z = load_golden("special.npz")
    boards = z["boards"]
    n = len(boards)
    st = oracle.status(boards)
    assert np.array_equal(st["legal_mask"], z["legal_mask"])
    assert np.array_equal(st["n_empty"], z["n_empty"])
    assert np.array_equal(st["highest_exp"], z["board_highest_exp"])
    assert np.array_equal(st["is_end"], z["is_end"][:, 0])
    assert np.array_equal(oracle.status(boards, max_tile_exp=11)["is_end"], z["is_end"][:, 1])
    for i in range(n):
        hi = int(z["board_highest_exp"][i])
        assert oracle.status(boards[i:i + 1], max_tile_exp=hi)["is_end"][0] == z["is_end"][i, 2]
    assert np.array_equal(oracle.encode_obs_u8(boards), z["obs"])
    for d in range(4):
        out, s, ch = oracle.move(boards, np.full(n, d, np.uint8))
        assert np.array_equal(out, z["move_boards"][:, d])
        assert np.array_equal(ch, z["move_changed"][:, d])
        assert np.array_equal(s[ch != 0], z["move_scores"][:, d][ch != 0])
        b = oracle.OracleBatch(n, auto_reset=False, illegal_move_reward=-1.0)
        b.boards[:] = boards
        o = b.step(np.full(n, d, np.uint8), forced_draws=z["words"][:, d])
        v = z["valid"][:, d] != 0
        assert np.array_equal(b.boards[v], z["out_boards"][:, d][v])
        assert np.array_equal(o["rewards"][v], z["rewards"][:, d][v])
        assert np.array_equal(o["dones"][v], z["dones"][:, d][v])
        assert np.array_equal(o["illegal"][v], z["illegal"][:, d][v])
        assert np.array_equal(o["highest_exp"][v], z["highest_exp"][:, d][v])


def test_rollouts_match_reference(golden_rollouts):
    for name, rec in golden_rollouts.items():
        cfg = rollout_cfg(rec)
        T = cfg.pop("T")
        b = oracle.OracleBatch(**cfg)
        assert np.array_equal(b.reset(), rec["init_boards"]), name
        for t in range(T):
            out = b.step(rec["actions"][t])
            out["ep_score"], out["ep_len"] = b.ep_score, b.ep_len
            for k in STEP_KEYS:
                assert np.array_equal(out[k], rec[k][t]), (name, t, k)
            d = rec["dones"][t] != 0
            for k in DONE_KEYS:
                assert np.array_equal(out[k][d], rec[k][t][d]), (name, t, k)


def test_sharding_invariance(golden_rollouts):
    """Two half-batches with env_id_base offsets reproduce the unsharded rollout."""
    rec = golden_rollouts["s42_legal_pen_base"]
    cfg = rollout_cfg(rec)
    T = min(cfg.pop("T"), 64)
    n = cfg.pop("n")
    base = cfg.pop("env_id_base")
    h = n // 2
    parts = [oracle.OracleBatch(h, env_id_base=base, **cfg), oracle.OracleBatch(n - h, env_id_base=base + h, **cfg)]
    init = np.concatenate([p.reset() for p in parts])
    assert np.array_equal(init, rec["init_boards"])
    for t in range(T):
        outs = [p.step(rec["actions"][t][i * h:(i * h + p.n)]) for i, p in enumerate(parts)]
        assert np.array_equal(np.concatenate([o["boards"] for o in outs]), rec["boards"][t])
        assert np.array_equal(np.concatenate([o["rewards"] for o in outs]), rec["rewards"][t])


def test_multithreaded_step_equals_single():
    a = oracle.OracleBatch(10000, seed=3, threads=1)
    b = oracle.OracleBatch(10000, seed=3, threads=5)
    a.reset(), b.reset()
    rng = np.random.default_rng(1)
    for _ in range(20):
        act = rng.integers(0, 4, 10000).astype(np.uint8)
        oa, ob = a.step(act), b.step(act)
        for k in oa:
            assert np.array_equal(oa[k], ob[k]), k
