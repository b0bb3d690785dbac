#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference env.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference (`/root/reference/env/envs/game2048_env.py`) is imported as is, with
`oracle/shim/gymnasium` standing in for the missing gymnasium package, and its RNG
replaced by `oracle.draws.InjectedDraws` so that add_tile (:166-176) places the tile
the frozen draw stream (include/g2048.h) selects.  Everything recorded here —
moved boards, rewards, terminated flags, highest, legality, one-hot observations —
is computed by the reference's own code.  Neither the C oracle nor the CUDA kernel is
involved in producing these files; tests compare both against them.

Outputs (all boards stored as uint8 exponents, value = 2**e, 0 = empty):
  shift_table.npz   exhaustive Game2048Env.shift over exponents 0..17 (18**4 rows)
  csv_transitions.npz  the 848 transitions of the reference's data/test_data.csv,
                    re-verified through reference step() with the forced draw
  rollouts.npz      injected-draw rollouts (auto-reset wrapper = SB3 DummyVecEnv
                    semantics) for seeds {0, 1, 42, 456}, several configurations
  special.npz       hand-made boards (dead, near-dead, exponents 16-17, max_tile)
                    stepped in all four directions + status/obs of each
"""
import csv
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("G2048_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
try:
    import gymnasium  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, REF)

from env.envs.game2048_env import Game2048Env, IllegalMove, stack  # noqa: E402  (the reference)
from oracle import draws  # noqa: E402


def to_exp(values):
    v = np.asarray(values, dtype=np.int64)
    e = np.zeros(v.shape, np.uint8)
    nz = v > 0
    e[nz] = np.round(np.log2(v[nz])).astype(np.uint8)
    assert np.array_equal(np.where(nz, np.int64(1) << e.astype(np.int64), 0), v)
    return e


def to_values(exps):
    e = np.asarray(exps).astype(np.int64)
    return np.where(e > 0, np.int64(1) << e, 0)


def ref_legal_mask(env):
    mask = 0
    for d in range(4):
        try:
            env.move(d, trial=True)
            mask |= 1 << d
        except IllegalMove:
            pass
    return mask


# --------------------------------------------------------------------------- shift
def make_shift_table():
    env = Game2048Env()
    rows = np.array(list(itertools.product(range(18), repeat=4)), dtype=np.uint8)
    out = np.zeros_like(rows)
    score = np.zeros(len(rows), np.uint32)
    for i, r in enumerate(rows):
        new, s = env.shift([int(x) for x in to_values(r)])
        out[i] = to_exp(new)
        score[i] = s
    np.savez_compressed(os.path.join(HERE, "shift_table.npz"), rows=rows, out=out, score=score)
    print("shift_table", rows.shape, int(score.max()))


# --------------------------------------------------------------------------- CSV
def make_csv_transitions():
    path = os.path.join(REF, "data", "test_data.csv")
    with open(path) as f:
        rd = csv.reader(f)
        next(rd)
        rows = [r for r in rd]
    n = len(rows)
    boards = np.zeros((n, 16), np.uint8)
    nxt = np.zeros((n, 16), np.uint8)
    actions = np.zeros(n, np.uint8)
    rewards = np.zeros(n, np.float32)
    done_csv = np.zeros(n, np.uint8)
    words = np.zeros((n, 4), np.uint32)
    highest = np.zeros(n, np.uint8)
    term_none = np.zeros(n, np.uint8)
    term_2048 = np.zeros(n, np.uint8)
    for i, r in enumerate(rows):
        b = np.array([int(float(x)) for x in r[0:16]], dtype=np.int64)
        a = int(float(r[16]))
        rew = float(r[17])
        nb = np.array([int(float(x)) for x in r[18:34]], dtype=np.int64)
        d = int(float(r[34]))
        # forced draw: the one cell where next differs from the reference-moved board
        env = Game2048Env()
        env.reset(seed=0)
        env.set_board(b.reshape(4, 4).copy())
        score = env.move(a)
        moved = env.get_board().reshape(-1).copy()
        assert score == rew, (i, score, rew)
        diff = np.flatnonzero(moved != nb)
        assert len(diff) == 1 and moved[diff[0]] == 0 and nb[diff[0]] in (2, 4), (i, diff)
        empties = np.flatnonzero(moved == 0)
        k = int(np.flatnonzero(empties == diff[0])[0])
        w = draws.word_for_spawn(len(empties), k, int(nb[diff[0]]))
        # replay through the reference's step() with that draw, for both max_tile settings
        for max_tile, dst in ((None, term_none), (2048, term_2048)):
            env = Game2048Env()
            env.reset(seed=0)
            env.set_max_tile(max_tile)
            env.set_board(b.reshape(4, 4).copy())
            inj = draws.InjectedDraws(env)
            env.np_random = inj
            inj.push(w)
            obs, reward, terminated, truncated, info = env.step(a)
            assert np.array_equal(env.get_board().reshape(-1), nb), i
            assert reward == rew and not info["illegal_move"] and not truncated
            dst[i] = terminated
            highest[i] = to_exp(info["highest"])
        boards[i], nxt[i], actions[i], rewards[i], done_csv[i] = to_exp(b), to_exp(nb), a, rew, d
        words[i] = (w, 0, 0, 0)
    np.savez_compressed(os.path.join(HERE, "csv_transitions.npz"), boards=boards, actions=actions,
                        rewards=rewards, next_boards=nxt, done_csv=done_csv, words=words,
                        highest_exp=highest, terminated_max_none=term_none,
                        terminated_max_2048=term_2048)
    print("csv_transitions", n, "done(max_tile=None)", int(term_none.sum()),
          "done(max_tile=2048)", int(term_2048.sum()), "max reward", float(rewards.max()))


# --------------------------------------------------------------------------- rollouts
def rollout(seed, n_envs, n_steps, env_id_base, policy, illegal_reward, max_tile, auto_reset,
            act_seed):
    """One injected-draw rollout of n_envs reference envs with DummyVecEnv auto-reset."""
    rng = np.random.default_rng(act_seed)
    env_ids = env_id_base + np.arange(n_envs, dtype=np.uint64)
    envs, injs = [], []
    w_reset = draws.draw_words(seed, env_ids, 0, 1)
    for i in range(n_envs):
        e = Game2048Env()
        e.set_illegal_move_reward(illegal_reward)
        e.set_max_tile(max_tile)
        inj = draws.InjectedDraws(e)
        e.np_random = inj
        inj.push(w_reset[i, 1], w_reset[i, 2])
        e.reset()                     # seed=None: keeps the injected np_random (:103)
        envs.append(e)
        injs.append(inj)
    init = np.stack([to_exp(e.get_board().reshape(-1)) for e in envs])
    T, B = n_steps, n_envs
    rec = dict(
        actions=np.zeros((T, B), np.uint8), boards=np.zeros((T, B, 16), np.uint8),
        rewards=np.zeros((T, B), np.float32), dones=np.zeros((T, B), np.uint8),
        illegal=np.zeros((T, B), np.uint8), highest_exp=np.zeros((T, B), np.uint8),
        legal_mask=np.zeros((T, B), np.uint8), terminal_boards=np.zeros((T, B, 16), np.uint8),
        final_score=np.zeros((T, B), np.uint32), final_len=np.zeros((T, B), np.uint32),
        ep_score=np.zeros((T, B), np.uint32), ep_len=np.zeros((T, B), np.uint32),
    )
    ep_len = np.zeros(B, np.int64)
    frozen = np.zeros(B, bool)        # without auto-reset a finished env keeps being stepped
    for t in range(T):
        w = draws.draw_words(seed, env_ids, t, 0)
        for i, (e, inj) in enumerate(zip(envs, injs)):
            mask = ref_legal_mask(e)
            if policy == "legal" and mask:
                a = int(rng.choice([d for d in range(4) if mask >> d & 1]))
            else:
                a = int(rng.integers(4))
            inj.words = [int(w[i, 0])]
            obs, reward, terminated, truncated, info = e.step(a)
            inj.words = []            # an illegal move consumes no draw (:91-95)
            assert truncated is False
            assert np.array_equal(obs, stack(e.get_board()))
            ep_len[i] += 1
            rec["actions"][t, i] = a
            rec["rewards"][t, i] = reward
            rec["dones"][t, i] = terminated
            rec["illegal"][t, i] = info["illegal_move"]
            rec["highest_exp"][t, i] = to_exp(info["highest"])
            if terminated:
                rec["terminal_boards"][t, i] = to_exp(e.get_board().reshape(-1))
                rec["final_score"][t, i] = int(e.score)
                rec["final_len"][t, i] = ep_len[i]
                if auto_reset:
                    inj.words = [int(w[i, 1]), int(w[i, 2])]
                    e.reset()
                    ep_len[i] = 0
            rec["boards"][t, i] = to_exp(e.get_board().reshape(-1))
            rec["legal_mask"][t, i] = ref_legal_mask(e)
            rec["ep_score"][t, i] = int(e.score)
            rec["ep_len"][t, i] = ep_len[i]
    rec["init_boards"] = init
    rec["meta"] = np.array([seed, n_envs, n_steps, env_id_base, 1 if policy == "legal" else 0,
                            0 if max_tile is None else int(np.log2(max_tile)),
                            1 if auto_reset else 0], dtype=np.int64)
    rec["illegal_reward"] = np.float32(illegal_reward)
    return rec


ROLLOUT_CONFIGS = [
    # name, seed, n_envs, n_steps, env_id_base, policy, illegal_reward, max_tile, auto_reset
    ("s0_random", 0, 96, 160, 0, "random", 0.0, None, True),
    ("s1_legal", 1, 96, 320, 0, "legal", 0.0, None, True),
    ("s42_legal_pen_base", 42, 64, 256, (1 << 33) + 12345, "legal", -1.0, None, True),
    ("s456_legal_max64", 456, 64, 200, 777, "legal", 0.0, 64, True),
    ("s0_random_noreset", 0, 64, 48, 65536, "random", -1.0, None, False),
    ("s1_legal_noreset", 1, 48, 400, 5, "legal", 0.0, None, False),
]


def make_rollouts():
    out = {}
    for name, seed, n_envs, n_steps, base, policy, ir, mt, ar in ROLLOUT_CONFIGS:
        rec = rollout(seed, n_envs, n_steps, base, policy, ir, mt, ar, act_seed=seed + 1000)
        for k, v in rec.items():
            out[f"{name}/{k}"] = v
        print("rollout", name, "steps", n_envs * n_steps, "dones", int(rec["dones"].sum()),
              "illegal", int(rec["illegal"].sum()), "max exp", int(rec["highest_exp"].max()))
    out["names"] = np.array([c[0] for c in ROLLOUT_CONFIGS])
    np.savez_compressed(os.path.join(HERE, "rollouts.npz"), **out)


# --------------------------------------------------------------------------- special boards
def special_boards():
    rng = np.random.default_rng(2048)
    B = []
    B.append([[2, 4, 8, 16], [4, 8, 16, 2], [8, 16, 2, 4], [16, 2, 4, 8]])          # dead
    B.append([[2, 4, 8, 16], [4, 8, 16, 2], [8, 16, 2, 4], [16, 2, 4, 0]])          # one hole
    B.append([[2, 2, 2, 2]] * 4)                                                    # all equal
    B.append([[0, 2, 0, 4], [2, 2, 8, 0], [2, 2, 2, 8], [2, 2, 4, 4]])              # ref test_move
    B.append([[2048, 0, 0, 0], [0] * 4, [0] * 4, [0] * 4])
    B.append([[1024, 1024, 0, 0], [0] * 4, [0] * 4, [0] * 4])
    B.append([[1024, 0, 0, 0], [1024, 0, 0, 0], [0] * 4, [0] * 4])
    B.append([[65536, 32768, 16384, 8192], [4096, 2048, 1024, 512], [256, 128, 64, 32], [16, 8, 4, 2]])
    B.append([[131072, 65536, 32768, 16384], [2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 0]])
    B.append([[32768, 32768, 0, 0], [16384, 16384, 16384, 16384], [0, 0, 0, 2], [4, 0, 0, 4]])
    B.append([[0, 0, 0, 2], [0] * 4, [0] * 4, [0] * 4])
    B.append([[2, 0, 0, 0], [0] * 4, [0] * 4, [0, 0, 4, 0]])
    B.append([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 4]])              # one pair left
    B.append([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [2, 2, 4, 2]])
    B.append([[4, 4, 4, 4], [4, 4, 4, 4], [8, 8, 8, 8], [8, 8, 8, 8]])
    B.append([[2, 2, 4, 8], [0, 0, 0, 0], [2, 2, 2, 2], [4, 0, 4, 0]])
    for _ in range(240):                                                            # random dense/sparse
        hi = int(rng.integers(2, 18))
        p_empty = float(rng.choice([0.0, 0.05, 0.2, 0.5, 0.8]))
        e = rng.integers(1, hi + 1, size=16)
        e[rng.random(16) < p_empty] = 0
        B.append(to_values(e).reshape(4, 4).tolist())
    return np.array(B, dtype=np.int64)


def make_special():
    boards = special_boards()
    n = len(boards)
    word_list = [0x00000000, 0xFFFFFFFF, 0x80000000, 0xE6666666, 0xE6666667, 0x12345678]
    out_boards = np.zeros((n, 4, 16), np.uint8)
    rewards = np.zeros((n, 4), np.float32)
    dones = np.zeros((n, 4), np.uint8)
    illegal = np.zeros((n, 4), np.uint8)
    highest = np.zeros((n, 4), np.uint8)
    valid = np.ones((n, 4), np.uint8)       # 0 where the reference itself asserts (:87)
    words = np.zeros((n, 4, 4), np.uint32)
    move_boards = np.zeros((n, 4, 16), np.uint8)
    move_scores = np.zeros((n, 4), np.uint32)
    move_changed = np.zeros((n, 4), np.uint8)
    legal_mask = np.zeros(n, np.uint8)
    n_empty = np.zeros(n, np.uint8)
    hi_exp = np.zeros(n, np.uint8)
    is_end = np.zeros((n, 3), np.uint8)     # max_tile None / 2048 / highest tile of the board
    obs = np.zeros((n, 16, 4, 4), np.uint8)
    for i, b in enumerate(boards):
        env = Game2048Env()
        env.reset(seed=0)
        env.set_board(b.copy())
        legal_mask[i] = ref_legal_mask(env)
        n_empty[i] = len(env.empties())
        hi_exp[i] = to_exp(env.highest())
        for j, mt in enumerate((None, 2048, int(env.highest()) if env.highest() > 0 else None)):
            env.set_max_tile(mt)
            is_end[i, j] = env.isend()
        o = stack(b)
        assert o.dtype == np.int64 and o.shape == (16, 4, 4)
        obs[i] = o
        for d in range(4):
            env = Game2048Env()
            env.reset(seed=0)
            env.set_board(b.copy())
            try:
                move_scores[i, d] = env.move(d)
                move_changed[i, d] = 1
            except IllegalMove:
                pass
            move_boards[i, d] = to_exp(env.get_board().reshape(-1))
            env = Game2048Env()
            env.reset(seed=0)
            env.set_illegal_move_reward(-1.0)
            env.set_board(b.copy())
            inj = draws.InjectedDraws(env)
            env.np_random = inj
            w = word_list[(i * 4 + d) % len(word_list)] ^ (i * 2654435761 & 0xFFFFFFFF)
            inj.push(w)
            words[i, d] = (w, 0, 0, 0)
            try:
                _, reward, terminated, truncated, info = env.step(d)
            except AssertionError:          # reference asserts score <= 2**16 (:87)
                valid[i, d] = 0
                continue
            out_boards[i, d] = to_exp(env.get_board().reshape(-1))
            rewards[i, d], dones[i, d] = reward, terminated
            illegal[i, d], highest[i, d] = info["illegal_move"], to_exp(info["highest"])
    np.savez_compressed(os.path.join(HERE, "special.npz"), boards=to_exp(boards.reshape(n, 16)),
                        words=words, out_boards=out_boards, rewards=rewards, dones=dones,
                        illegal=illegal, highest_exp=highest, valid=valid,
                        move_boards=move_boards, move_scores=move_scores,
                        move_changed=move_changed, legal_mask=legal_mask, n_empty=n_empty,
                        board_highest_exp=hi_exp, is_end=is_end, obs=obs)
    print("special", n, "invalid(ref asserts)", int((valid == 0).sum()))


if __name__ == "__main__":
    make_shift_table()
    make_csv_transitions()
    make_special()
    make_rollouts()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
