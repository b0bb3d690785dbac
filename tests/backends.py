"""Test backends with one interface: the C oracle, the host simulation of the device
header, and (on a GPU box) the product through its C ABI.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_sim = None


def sim_lib():
    global _sim
    if _sim is None:
        d = os.path.join(HERE, "host_sim")
        so, src = os.path.join(d, "libsim.so"), os.path.join(d, "sim.cpp")
        hdr = os.path.join(ROOT, "gym-2048_b200", "csrc", "g2048_device.cuh")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-o", so, src])
        _sim = C.CDLL(so)
    return _sim


class SimBatch(oracle.OracleBatch):
    """Same state/arguments as OracleBatch, stepped by the host-compiled device logic."""

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        sim_lib().sim_reset(_p(self.boards), _p(m), C.c_uint64(self.n), C.c_uint64(self.env_id_base),
                            C.c_uint64(self.seed), C.c_uint64(self.reset_index))
        self.reset_index += 1
        sel = slice(None) if mask is None else (m != 0)
        self.ep_score[sel] = 0
        self.ep_len[sel] = 0
        self.ep_return[sel] = 0
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
        a = oracle.StepArgs(_p(self.boards), _p(actions), _p(out["rewards"]), _p(out["dones"]),
                            _p(out["illegal"]), _p(out["highest_exp"]), _p(out["legal_mask"]),
                            _p(out["terminal_boards"]), _p(self.ep_score), _p(self.ep_len),
                            _p(out["final_score"]), _p(out["final_len"]), _p(fd), None,
                            n, self.env_id_base, self.seed, self.step_index,
                            self.illegal_move_reward, self.max_tile_exp, self.flags, None,
                            _p(self.ep_return), _p(out["final_return"]))
        assert sim_lib().sim_step(C.byref(a)) == 0
        self.step_index += 1
        out["boards"] = self.boards
        return out


class SimOps:
    name = "sim"
    Batch = SimBatch

    @staticmethod
    def move(boards, directions):
        b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
        d = np.ascontiguousarray(directions, dtype=np.uint8)
        out, scores, changed = np.empty_like(b), np.zeros(len(b), np.uint32), np.zeros(len(b), np.uint8)
        sim_lib().sim_move(_p(b), _p(out), _p(d), _p(scores), _p(changed), C.c_uint64(len(b)))
        return out, scores, changed

    @staticmethod
    def add_tile(boards, env_id_base, seed, step_index):
        b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16).copy()
        sim_lib().sim_add_tile(_p(b), C.c_uint64(len(b)), C.c_uint64(env_id_base), C.c_uint64(seed),
                               C.c_uint64(step_index))
        return b

    @staticmethod
    def status(boards, max_tile_exp=0):
        b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
        o = [np.zeros(len(b), np.uint8) for _ in range(4)]
        sim_lib().sim_status(_p(b), _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]), C.c_uint32(max_tile_exp),
                             C.c_uint64(len(b)))
        return dict(legal_mask=o[0], highest_exp=o[1], n_empty=o[2], is_end=o[3])

    @staticmethod
    def philox(ctr, k0, k1):
        c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
        out = np.empty_like(c)
        sim_lib().sim_philox(_p(c), C.c_uint32(k0), C.c_uint32(k1), _p(out), C.c_uint64(len(c)))
        return out

    @staticmethod
    def philox2x32(ctr, key):
        c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 2)
        out = np.empty_like(c)
        sim_lib().sim_philox2x32(_p(c), C.c_uint32(key), _p(out), C.c_uint64(len(c)))
        return out

    @staticmethod
    def draw_words(n, env_id_base, seed, index, tag, form=1):
        out = np.empty((n, 4), np.uint32)
        sim_lib().sim_draw_words(_p(out), C.c_uint64(n), C.c_uint64(env_id_base), C.c_uint64(seed), C.c_uint64(index),
                                 C.c_uint32(tag), C.c_int(form))
        return out


class OracleOps:
    name = "oracle"
    Batch = oracle.OracleBatch
    move = staticmethod(oracle.move)
    add_tile = staticmethod(oracle.add_tile)
    status = staticmethod(oracle.status)
    philox = staticmethod(oracle.philox)
    philox2x32 = staticmethod(oracle.philox2x32)
    draw_words = staticmethod(oracle.draw_words)


# ---------------------------------------------------------------------------------------
# The product on a GPU, through libg2048.so's C ABI (gym_2048_b200.BatchedGame2048).
# ---------------------------------------------------------------------------------------
class GpuBatch:
    """OracleBatch-shaped view of BatchedGame2048 (numpy in / numpy out)."""

    def __init__(self, n, seed=0, env_id_base=0, illegal_move_reward=0.0, max_tile_exp=0, auto_reset=True,
                 threads=1):
        import gym_2048_b200 as g
        self.n = n
        self.g = g.BatchedGame2048(n, seed=seed, env_id_base=env_id_base,
                                   illegal_move_reward=illegal_move_reward,
                                   max_tile=(1 << max_tile_exp) if max_tile_exp else None, auto_reset=auto_reset)
        self._boards = _BoardsView(self.g)

    @property
    def boards(self):
        return self._boards

    @property
    def ep_score(self):
        return self.g.ep_score.cpu().numpy().astype(np.uint32)

    @property
    def ep_len(self):
        return self.g.ep_len.cpu().numpy().astype(np.uint32)

    @property
    def ep_return(self):
        return self.g.ep_return.cpu().numpy()

    def reset(self, mask=None):
        import torch
        m = None if mask is None else torch.as_tensor(np.ascontiguousarray(mask, dtype=np.uint8))
        return self.g.reset(mask=m).cpu().numpy()

    def step(self, actions, forced_draws=None):
        import torch
        fd = None if forced_draws is None else torch.from_numpy(np.ascontiguousarray(forced_draws, dtype=np.uint32))
        r = self.g.step(torch.from_numpy(np.ascontiguousarray(actions, dtype=np.uint8)), forced_draws=fd)
        c = lambda t, dt: t.cpu().numpy().astype(dt)       # noqa: E731
        return dict(boards=c(r.boards, np.uint8), rewards=c(r.rewards, np.float32), dones=c(r.dones, np.uint8),
                    illegal=c(r.illegal, np.uint8), highest_exp=c(r.highest_exp, np.uint8),
                    legal_mask=c(r.legal_mask, np.uint8), terminal_boards=c(r.terminal_boards, np.uint8),
                    final_score=c(r.final_score, np.uint32), final_len=c(r.final_len, np.uint32),
                    final_return=c(r.final_return, np.float32))


class _BoardsView:
    """`b.boards[:] = x` / `b.boards[i] = x` / np.array_equal(b.boards, y) against device state."""

    def __init__(self, g):
        self.g = g

    def __setitem__(self, key, value):
        import torch
        cur = self.g.boards.cpu().numpy()
        cur[key] = value
        self.g.set_boards(torch.from_numpy(cur))

    def __getitem__(self, key):
        return self.g.boards.cpu().numpy()[key]

    def __array__(self, dtype=None, copy=None):
        a = self.g.boards.cpu().numpy()
        return a if dtype is None else a.astype(dtype)


class GpuOps:
    name = "gpu"
    Batch = GpuBatch

    @staticmethod
    def _game(boards):
        import gym_2048_b200 as g
        import torch
        b = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, 16)
        game = g.BatchedGame2048(len(b), outputs=())
        game.set_boards(torch.from_numpy(b))
        return game

    @staticmethod
    def move(boards, directions):
        import torch
        game = GpuOps._game(boards)
        s, ch = game.move(torch.from_numpy(np.ascontiguousarray(directions, dtype=np.uint8)))
        return game.boards.cpu().numpy(), s.cpu().numpy().astype(np.uint32), ch.cpu().numpy().astype(np.uint8)

    @staticmethod
    def status(boards, max_tile_exp=0):
        game = GpuOps._game(boards)
        game.max_tile_exp = max_tile_exp
        st = game.status()
        return {k: v.cpu().numpy().astype(np.uint8) for k, v in st.items()}

    @staticmethod
    def add_tile(boards, env_id_base, seed, step_index):
        import ctypes as C
        import torch
        game = GpuOps._game(boards)
        from gym_2048_b200._lib import check
        check(game.lib.g2048_add_tile(C.c_void_p(game.boards.data_ptr()), game.num_envs, env_id_base, seed,
                                      step_index, None))
        torch.cuda.synchronize()
        return game.boards.cpu().numpy()

    @staticmethod
    def philox(ctr, k0, k1):
        import ctypes as C
        import torch
        import gym_2048_b200 as g
        from gym_2048_b200._lib import check
        c = torch.from_numpy(np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)).cuda()
        out = torch.empty_like(c)
        check(g._lib.lib().g2048_philox(C.c_void_p(c.data_ptr()), k0, k1, C.c_void_p(out.data_ptr()), c.shape[0],
                                        None))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    @staticmethod
    def philox2x32(ctr, key):
        import ctypes as C
        import torch
        import gym_2048_b200 as g
        from gym_2048_b200._lib import check
        c = torch.from_numpy(np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 2)).cuda()
        out = torch.empty_like(c)
        check(g._lib.lib().g2048_philox2x32(C.c_void_p(c.data_ptr()), key, C.c_void_p(out.data_ptr()), c.shape[0], None))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    @staticmethod
    def draw_words(n, env_id_base, seed, index, tag):
        import ctypes as C
        import torch
        import gym_2048_b200 as g
        from gym_2048_b200._lib import check
        out = torch.empty((n, 4), dtype=torch.uint32, device="cuda")
        check(g._lib.lib().g2048_draw_words(C.c_void_p(out.data_ptr()), n, env_id_base, seed, index, tag, None))
        torch.cuda.synchronize()
        return out.cpu().numpy()
