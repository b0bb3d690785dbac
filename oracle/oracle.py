"""ctypes front-end of libg2048_oracle.so (the C restatement) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libg2048_oracle.so")
FLAG_AUTO_RESET = 1


class StepArgs(C.Structure):
    """Mirror of G2048StepArgs (include/g2048.h)."""
    _fields_ = [
        ("boards", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p),
        ("dones", C.c_void_p), ("illegal", C.c_void_p), ("highest_exp", C.c_void_p),
        ("legal_mask", C.c_void_p), ("terminal_boards", C.c_void_p),
        ("ep_score", C.c_void_p), ("ep_len", C.c_void_p),
        ("final_score", C.c_void_p), ("final_len", C.c_void_p),
        ("forced_draws", C.c_void_p), ("step_counter", C.c_void_p),
        ("n", C.c_uint64), ("env_id_base", C.c_uint64), ("seed", C.c_uint64),
        ("step_index", C.c_uint64),
        ("illegal_move_reward", C.c_float), ("max_tile_exp", C.c_uint32),
        ("flags", C.c_uint32), ("boards_out", C.c_void_p),
        ("ep_return", C.c_void_p), ("final_return", C.c_void_p),
        ("boards_nibble", C.c_void_p), ("nibble_overflow", C.c_void_p),
        ("chain", C.c_void_p),          # a launch-scheduling device buffer of the CUDA library; the oracle ignores it
    ]


def build(force=False):
    src = os.path.join(_HERE, "g2048_oracle.c")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "g2048.h")      # the argument structs come from the public header
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.g2048_oracle_shift.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def values_to_exp(values):
    """int tile values (0,2,4,...) -> uint8 exponents; shape preserved."""
    v = np.asarray(values, dtype=np.int64)
    e = np.zeros(v.shape, dtype=np.uint8)
    nz = v > 0
    e[nz] = np.round(np.log2(v[nz])).astype(np.uint8)
    assert np.array_equal(np.where(nz, np.int64(1) << e.astype(np.int64), 0), v), "not powers of two"
    return e


def exp_to_values(exps):
    e = np.asarray(exps).astype(np.int64)
    return np.where(e > 0, np.int64(1) << e, 0)


class OracleBatch:
    """Batched env state stepped by the C oracle; same argument meaning as g2048_step."""

    def __init__(self, n, seed=0, env_id_base=0, illegal_move_reward=0.0, max_tile_exp=0,
                 auto_reset=True, threads=1):
        self.n, self.seed, self.env_id_base = int(n), int(seed), int(env_id_base)
        self.illegal_move_reward, self.max_tile_exp = float(illegal_move_reward), int(max_tile_exp)
        self.flags = FLAG_AUTO_RESET if auto_reset else 0
        self.threads = threads
        self.boards = np.zeros((self.n, 16), np.uint8)
        self.ep_score = np.zeros(self.n, np.uint32)
        self.ep_len = np.zeros(self.n, np.uint32)
        self.ep_return = np.zeros(self.n, np.float32)
        self.step_index = 0
        self.reset_index = 0

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = lib().g2048_oracle_reset(_p(self.boards), _p(m), C.c_uint64(self.n),
                                      C.c_uint64(self.env_id_base), C.c_uint64(self.seed),
                                      C.c_uint64(self.reset_index))
        assert rc == 0
        self.reset_index += 1
        if mask is None:
            self.ep_score[:] = 0
            self.ep_len[:] = 0
            self.ep_return[:] = 0
        else:
            self.ep_score[m != 0] = 0
            self.ep_len[m != 0] = 0
            self.ep_return[m != 0] = 0
        return self.boards

    def step(self, actions, forced_draws=None):
        n = self.n
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        out = dict(
            rewards=np.zeros(n, np.float32), dones=np.zeros(n, np.uint8),
            illegal=np.zeros(n, np.uint8), highest_exp=np.zeros(n, np.uint8),
            legal_mask=np.zeros(n, np.uint8), terminal_boards=np.zeros((n, 16), np.uint8),
            final_score=np.zeros(n, np.uint32), final_len=np.zeros(n, np.uint32),
            final_return=np.zeros(n, np.float32),
        )
        fd = None if forced_draws is None else np.ascontiguousarray(forced_draws, dtype=np.uint32)
        a = StepArgs(_p(self.boards), _p(actions), _p(out["rewards"]), _p(out["dones"]),
                     _p(out["illegal"]), _p(out["highest_exp"]), _p(out["legal_mask"]),
                     _p(out["terminal_boards"]), _p(self.ep_score), _p(self.ep_len),
                     _p(out["final_score"]), _p(out["final_len"]), _p(fd), None,
                     n, self.env_id_base, self.seed, self.step_index,
                     self.illegal_move_reward, self.max_tile_exp, self.flags, None,
                     _p(self.ep_return), _p(out["final_return"]))
        if self.threads > 1:
            rc = lib().g2048_oracle_step_mt(C.byref(a), C.c_int(self.threads))
        else:
            rc = lib().g2048_oracle_step(C.byref(a))
        assert rc == 0
        self.step_index += 1
        out["boards"] = self.boards
        return out


    def step_many(self, actions):
        """K consecutive step() calls (the definition of g2048_step_many): per-step outputs stacked [K, ...]."""
        outs = [self.step(a) for a in np.asarray(actions)]
        stacked = {k: np.stack([o[k] for o in outs]) for k in ("rewards", "dones", "illegal")}
        return stacked


def add_tile(boards, env_id_base, seed, step_index):
    b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16).copy()
    rc = lib().g2048_oracle_add_tile(_p(b), C.c_uint64(len(b)), C.c_uint64(env_id_base), C.c_uint64(seed),
                                     C.c_uint64(step_index))
    assert rc == 0
    return b


def move(boards, directions):
    b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
    d = np.ascontiguousarray(directions, dtype=np.uint8)
    out = np.empty_like(b)
    scores = np.zeros(len(b), np.uint32)
    changed = np.zeros(len(b), np.uint8)
    rc = lib().g2048_oracle_move(_p(b), _p(out), _p(d), _p(scores), _p(changed), C.c_uint64(len(b)))
    assert rc == 0
    return out, scores, changed


def status(boards, max_tile_exp=0):
    b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
    n = len(b)
    o = [np.zeros(n, np.uint8) for _ in range(4)]
    rc = lib().g2048_oracle_status(_p(b), _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]),
                                   C.c_uint32(max_tile_exp), C.c_uint64(n))
    assert rc == 0
    return dict(legal_mask=o[0], highest_exp=o[1], n_empty=o[2], is_end=o[3])


def encode_obs_u8(boards):
    b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
    obs = np.empty((len(b), 16, 4, 4), np.uint8)
    rc = lib().g2048_oracle_encode_obs_u8(_p(b), _p(obs), C.c_uint64(len(b)))
    assert rc == 0
    return obs


def philox(ctr, key0, key1):
    c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
    out = np.empty_like(c)
    rc = lib().g2048_oracle_philox(_p(c), C.c_uint32(key0), C.c_uint32(key1), _p(out), C.c_uint64(len(c)))
    assert rc == 0
    return out


def philox2x32(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 2)
    out = np.empty_like(c)
    rc = lib().g2048_oracle_philox2x32(_p(c), C.c_uint32(key), _p(out), C.c_uint64(len(c)))
    assert rc == 0
    return out


def draw_words(n, env_id_base, seed, index, tag):
    out = np.empty((n, 4), np.uint32)
    rc = lib().g2048_oracle_draw_words(_p(out), C.c_uint64(n), C.c_uint64(env_id_base), C.c_uint64(seed),
                                       C.c_uint64(index), C.c_uint32(tag))
    assert rc == 0
    return out


def shift(row_exps):
    r = (C.c_uint8 * 4)(*[int(x) for x in row_exps])
    o = (C.c_uint8 * 4)()
    s = lib().g2048_oracle_shift(r, o)
    return list(o), int(s)


def sample_actions(legal_mask, n, env_id_base, seed, step_index):
    m = None if legal_mask is None else np.ascontiguousarray(legal_mask, dtype=np.uint8)
    out = np.zeros(n, np.uint8)
    rc = lib().g2048_oracle_sample_actions(_p(m), _p(out), C.c_uint64(n), C.c_uint64(env_id_base),
                                           C.c_uint64(seed), C.c_uint64(step_index))
    assert rc == 0
    return out


def symmetry(boards, actions, hflip, k):
    b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
    a = None if actions is None else np.ascontiguousarray(actions, dtype=np.uint8)
    ob = np.empty_like(b)
    oa = None if a is None else np.empty_like(a)
    rc = lib().g2048_oracle_symmetry(_p(b), _p(ob), _p(a), _p(oa), C.c_uint64(len(b)), C.c_int(hflip), C.c_int(k))
    assert rc == 0
    return ob, oa


def augment(boards, next_boards, actions, rewards, dones):
    b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
    nb = np.ascontiguousarray(next_boards, dtype=np.uint8).reshape(-1, 16)
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    r = np.ascontiguousarray(rewards, dtype=np.float32)
    d = np.ascontiguousarray(dones, dtype=np.uint8)
    n = len(b)
    o = dict(boards=np.empty((8 * n, 16), np.uint8), next_boards=np.empty((8 * n, 16), np.uint8),
             actions=np.empty(8 * n, np.uint8), rewards=np.empty(8 * n, np.float32), dones=np.empty(8 * n, np.uint8))
    rc = lib().g2048_oracle_augment(_p(b), _p(nb), _p(a), _p(r), _p(d), C.c_uint64(n), _p(o["boards"]),
                                    _p(o["next_boards"]), _p(o["actions"]), _p(o["rewards"]), _p(o["dones"]))
    assert rc == 0
    return o


def discounted_return(rewards, dones, gamma=0.9):
    r = np.ascontiguousarray(rewards, dtype=np.float32)
    d = np.ascontiguousarray(dones, dtype=np.uint8)
    out = np.zeros(len(r), np.float64)
    rc = lib().g2048_oracle_discounted_return(_p(r), _p(d), _p(out), C.c_uint64(len(r)), C.c_double(gamma))
    assert rc == 0
    return out
