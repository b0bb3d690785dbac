"""Game2048Env — the reference's single-env gymnasium class, executed on the GPU.

Same constructor, attributes, method names, argument meaning and error behaviour as
`/root/reference/env/envs/game2048_env.py` (class :34-288, `stack` :17-32, `IllegalMove`
:14-15) so the reference's callers (train.py:150-165,183-184; gather_training_data.py
:91,141-145,191; env/envs/test_game2048_env.py) run unchanged.  Every game rule is evaluated
by the CUDA kernels of libg2048.so on a batch of one board; this class only marshals the
4x4 int64 `Matrix` to and from the device.  What differs from the reference, by design:
the spawn RNG is the counter-based Philox stream of include/g2048.h keyed by `seed`
(the reference draws from numpy's PCG64, which no reference test pins — SURVEY.md §8c).
"""
import ctypes as C
import os
import sys
from io import StringIO

import numpy as np
import torch

from . import _lib
from ._lib import check
from .batched import tile_to_exp

try:  # gymnasium is optional: the class works without it and registers itself when present
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _EnvBase = _gym.Env
except ImportError:  # pragma: no cover - depends on the image
    _gym = None
    _spaces = None
    _EnvBase = object


class IllegalMove(Exception):
    pass


class _Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = n, (), np.dtype(np.int64)
        self._rng = np.random.default_rng()

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

    __contains__ = contains


class _Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high = np.full(shape, low, dtype=dtype), np.full(shape, high, dtype=dtype)
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains


def _discrete(n):
    return _spaces.Discrete(n) if _spaces is not None else _Discrete(n)


def _box(low, high, shape, dtype):
    return _spaces.Box(low, high, shape, dtype=dtype) if _spaces is not None else _Box(low, high, shape, dtype)


class _Device:
    """Per-process scratch for one-board launches (device buffers + the loaded library)."""
    _inst = None

    def __init__(self):
        if not torch.cuda.is_available():
            raise _lib.G2048Error("Game2048Env needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.lib()
        self.dev = torch.device("cuda", int(os.environ.get("G2048_DEVICE", torch.cuda.current_device())))
        d = self.dev
        self.values = torch.zeros(16, dtype=torch.int64, device=d)
        self.obs = torch.zeros((16, 4, 4), dtype=torch.int64, device=d)
        self.boards = torch.zeros(64, dtype=torch.uint8, device=d)          # 4 boards (one per direction)
        self.u8 = torch.zeros(64, dtype=torch.uint8, device=d)              # small byte outputs
        self.i32 = torch.zeros(8, dtype=torch.int32, device=d)
        self.f32 = torch.zeros(1, dtype=torch.float32, device=d)

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)


def _ptr(t, offset=0):
    return C.c_void_p(t.data_ptr() + offset)


def stack(flat, layers=15):
    """[4,4] tile values -> [layers+1,4,4] one-hot (reference :17-32): channel 0 = empty,
    channel k = (cell == 2**k).  Computed by g2048_encode_obs for the default 15 layers."""
    flat = np.asarray(flat)
    if layers != 15 or flat.shape != (4, 4):
        raise ValueError("stack() supports the reference's 4x4 board with layers=15")
    d = _Device.get()
    with torch.cuda.device(d.dev):
        d.values.copy_(torch.from_numpy(np.ascontiguousarray(flat, dtype=np.int64).reshape(16)))
        # cells that are not 0 / a power of two light no channel in the reference either
        check(d.lib.g2048_exp_from_values(_ptr(d.values), _ptr(d.boards), 16, None, d.stream()))
        bad = (d.values != 0) & (d.boards[:16] == 0)
        d.boards[:16].masked_fill_(bad, 63)
        check(d.lib.g2048_encode_obs(_ptr(d.boards), _ptr(d.obs), _lib.OBS_I64, 1, d.stream()))
        return d.obs.cpu().numpy().astype(int)


class Game2048Env(_EnvBase):
    metadata = {'render_modes': ['ansi', 'human', 'rgb_array'], 'render_fps': 4}

    def __init__(self, render_mode=None):
        self.size = 4
        self.w = self.size
        self.h = self.size
        self.squares = self.size * self.size
        self.score = 0
        self.action_space = _discrete(4)                                                   # :49
        self.observation_space = _box(0, 1, (self.squares, self.w, self.h), int)           # :51-52
        self.set_illegal_move_reward(0.)
        self.set_max_tile(None)
        self.grid_size = 70
        self.render_mode = render_mode
        self.Matrix = np.zeros((self.h, self.w), int)
        # draw stream position: (seed, reset_index, step_index) — see include/g2048.h
        self._seed = int.from_bytes(os.urandom(8), "little")
        self._step_index = 0
        self._reset_index = 0
        self._np_random_compat = None

    # -- knobs (:61-73) ------------------------------------------------------------------
    def set_illegal_move_reward(self, reward):
        self.illegal_move_reward = reward
        self.reward_range = (self.illegal_move_reward, float(2**self.squares))

    def set_max_tile(self, max_tile):
        assert max_tile is None or isinstance(max_tile, int)
        self.max_tile = max_tile

    @property
    def np_random(self):
        """Present for API compatibility; tile spawns use the Philox stream, not this."""
        if self._np_random_compat is None:
            self._np_random_compat = np.random.default_rng(self._seed & 0xFFFFFFFF)
        return self._np_random_compat

    @np_random.setter
    def np_random(self, value):
        self._np_random_compat = value

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass

    # -- device marshalling --------------------------------------------------------------
    def _upload(self, d, copies=1):
        v = torch.from_numpy(np.ascontiguousarray(self.Matrix, dtype=np.int64).reshape(16))
        d.values.copy_(v)
        bad = d.i32[7:8]
        bad.zero_()
        check(d.lib.g2048_exp_from_values(_ptr(d.values), _ptr(d.boards), 16, _ptr(bad), d.stream()))
        if int(bad.item()):
            raise ValueError("board holds a cell that is not 0 or a power of two: %s" % (self.Matrix,))
        for k in range(1, copies):
            d.boards[16 * k:16 * k + 16].copy_(d.boards[:16])

    def _download(self, d, offset=0):
        check(d.lib.g2048_values_from_exp(_ptr(d.boards, offset), _ptr(d.values), 16, d.stream()))
        # write through the caller's array: set_board() aliases it (:286-288)
        self.Matrix[...] = d.values.cpu().numpy().reshape(4, 4)

    # -- gymnasium interface -------------------------------------------------------------
    def step(self, action):
        """(:76-100) move, add a tile, detect the end.  An illegal move terminates (:91-95)."""
        d = _Device.get()
        info = {'illegal_move': False}
        with torch.cuda.device(d.dev):
            self._upload(d)
            d.u8[0] = int(action) & 3
            a = _lib.StepArgs(_ptr(d.boards), _ptr(d.u8, 0), _ptr(d.f32), _ptr(d.u8, 16), _ptr(d.u8, 17),
                              _ptr(d.u8, 18), None, None, None, None, None, None, None, None,
                              1, 0, self._seed, self._step_index,
                              float(self.illegal_move_reward), tile_to_exp(self.max_tile), 0)
            check(d.lib.g2048_step(C.byref(a), d.stream()))
            self._step_index += 1
            self._download(d)
            flags = d.u8[16:19].cpu().numpy()
            reward = float(d.f32.item())
        terminated = bool(flags[0])
        if flags[1]:
            info['illegal_move'] = True
            reward = self.illegal_move_reward
        else:
            assert reward <= 2**(self.w * self.h)                                          # :87
            self.score += reward
        info['highest'] = self.highest()
        return stack(self.Matrix), reward, terminated, False, info

    def reset(self, seed=None, options=None):
        """(:102-111) zero board, score 0, two tiles.  `seed` re-keys the draw stream."""
        if _gym is not None:
            super().reset(seed=seed)
        if seed is not None:
            self._seed = int(seed) & (2**64 - 1)
            self._step_index = 0
            self._reset_index = 0
        d = _Device.get()
        with torch.cuda.device(d.dev):
            check(d.lib.g2048_reset(_ptr(d.boards), None, 1, 0, self._seed, self._reset_index, d.stream()))
            self._reset_index += 1
            self.Matrix = np.zeros((self.h, self.w), int)
            self._download(d)
        self.score = 0
        return stack(self.Matrix), {}

    def render(self, mode=None):
        if mode is None:
            mode = self.render_mode or 'human'
        if mode == 'rgb_array':
            raise NotImplementedError("rgb_array rendering (reference :116-154, Pillow UI) is out of scope")
        outfile = StringIO() if mode == 'ansi' else sys.stdout
        outfile.write('Score: {}\nHighest: {}\n{}\n'.format(self.score, self.highest(), np.array(self.Matrix)))
        return outfile

    # -- game API ------------------------------------------------------------------------
    def add_tile(self):
        """(:166-176) spawn a 2 (P=0.9) or 4 on a uniformly chosen empty cell."""
        assert (np.asarray(self.Matrix) == 0).any(), "No empty cell found"                 # :176
        d = _Device.get()
        with torch.cuda.device(d.dev):
            self._upload(d)
            check(d.lib.g2048_add_tile(_ptr(d.boards), 1, 0, self._seed, self._step_index, d.stream()))
            self._step_index += 1
            self._download(d)

    def get(self, x, y):
        return self.Matrix[x, y]

    def set(self, x, y, val):
        self.Matrix[x, y] = val

    def empties(self):
        return np.argwhere(self.Matrix == 0)

    def _status(self):
        d = _Device.get()
        with torch.cuda.device(d.dev):
            self._upload(d)
            check(d.lib.g2048_status(_ptr(d.boards), _ptr(d.u8, 0), _ptr(d.u8, 1), _ptr(d.u8, 2), _ptr(d.u8, 3),
                                     tile_to_exp(self.max_tile), 1, d.stream()))
            return d.u8[:4].cpu().numpy()        # legal mask, highest exp, empties, isend

    def highest(self):
        """(:190-192) the highest tile on the board (numpy int, like np.max)."""
        e = int(self._status()[1])
        return np.int64((1 << e) if e else 0)

    def legal_actions(self):
        """Bit mask of legal moves (bit d <=> move(d, trial=True) does not raise)."""
        return int(self._status()[0])

    def move(self, direction, trial=False):
        """(:194-241) slide/merge toward 0=up 1=right 2=down 3=left; returns the score, raises
        IllegalMove when nothing moves.  trial=True leaves the board untouched."""
        d = _Device.get()
        with torch.cuda.device(d.dev):
            self._upload(d)
            d.u8[0] = int(direction) & 3
            check(d.lib.g2048_move(_ptr(d.boards), None if trial else _ptr(d.boards), _ptr(d.u8, 0),
                                   _ptr(d.i32), _ptr(d.u8, 16), 1, d.stream()))
            changed = bool(d.u8[16].item())
            score = int(d.i32[0].item())
            if changed and not trial:
                self._download(d)
        if not changed:
            raise IllegalMove
        return score

    def shift(self, row):
        """(:243-260) compact and combine one line toward index 0: ([4 values], score)."""
        row = [int(v) for v in row]
        assert len(row) == self.size
        d = _Device.get()
        with torch.cuda.device(d.dev):
            vals = torch.zeros(16, dtype=torch.int64)
            vals[:4] = torch.tensor(row, dtype=torch.int64)          # the line as row 0, moved Left
            d.values.copy_(vals)
            bad = d.i32[7:8]
            bad.zero_()
            check(d.lib.g2048_exp_from_values(_ptr(d.values), _ptr(d.boards), 16, _ptr(bad), d.stream()))
            if int(bad.item()):
                raise ValueError("shift() needs tile values that are 0 or powers of two")
            d.u8[0] = 3
            check(d.lib.g2048_move(_ptr(d.boards), _ptr(d.boards), _ptr(d.u8, 0), _ptr(d.i32), None, 1,
                                   d.stream()))
            check(d.lib.g2048_values_from_exp(_ptr(d.boards), _ptr(d.values), 16, d.stream()))
            out = d.values[:4].cpu().tolist()
            score = int(d.i32[0].item())
        return (out, score)

    def isend(self):
        """(:262-280) max_tile reached, or no empty cell and no legal move."""
        return bool(self._status()[3])

    def get_board(self):
        return self.Matrix

    def set_board(self, new_board):
        self.Matrix = new_board


def register(force=False):
    """Register '2048-v0' with gymnasium when it is installed (reference env/__init__.py:1-6, which does so
    when the package is imported; `import gym_2048_b200` calls this too unless G2048_NO_REGISTER=1).
    An id already taken — the reference package imported first — is left alone unless `force`."""
    if _gym is None:
        return False
    from gymnasium.envs.registration import register as _register, registry
    if force or '2048-v0' not in registry:
        _register(id='2048-v0', entry_point='gym_2048_b200.env:Game2048Env')
    return True
