"""Game2048Env — the reference's single-env gymnasium class, executed on the GPU.

Same constructor, attributes, method names, argument meaning and error behaviour as
`/root/reference/env/envs/game2048_env.py` (class :34-288, `stack` :17-32, `IllegalMove`
:14-15) so the reference's callers (train.py:150-165,183-184; gather_training_data.py
:91,141-145,191; env/envs/test_game2048_env.py) run unchanged.  Every game rule is evaluated
by libg2048.so's g2048_one kernel (one launch + one stream synchronisation per method call, the
Matrix travelling through a pinned zero-copy block); this class only marshals the 4x4 `Matrix`.  What differs from the reference, by design:
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
from .batched import _raw_stream, tile_to_exp

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
    """Per-process state of the single-env path: the loaded library and ONE packed in/out block (G2048OneIO) in
    pinned host memory that the kernel reads and writes in place (zero copy).  Every method of Game2048Env is one
    g2048_one call = one kernel launch + one stream synchronisation; nothing else touches the device."""
    _inst = None

    def __init__(self):
        if not torch.cuda.is_available():
            raise _lib.G2048Error("Game2048Env needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.lib()
        self.index = int(os.environ.get("G2048_DEVICE", torch.cuda.current_device()))
        self.dev = torch.device("cuda", self.index)
        self.buf = torch.zeros(C.sizeof(_lib.OneIO) + 64, dtype=torch.uint8).pin_memory()
        base = (self.buf.data_ptr() + 63) // 64 * 64
        self.io = _lib.OneIO.from_address(base)
        self.ptr = C.c_void_p(base)
        self.values = np.ctypeslib.as_array(self.io.values)                    # int64 [16]: the Matrix, in/out
        self.obs = np.ctypeslib.as_array(self.io.obs).reshape(16, 4, 4)        # int64 [16,4,4]: stack(Matrix), out

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def call(self, op, matrix, action=0, trial=False, seed=0, index=0, illegal_move_reward=0.0, max_tile_exp=0,
             need_valid=True):
        """values <- matrix; run `op`; returns the io block (results valid until the next call)."""
        if matrix is not None:
            self.values[:] = np.asarray(matrix).reshape(16)
        args = (self.ptr, op, int(action) & 3, 1 if trial else 0, seed, index, float(illegal_move_reward),
                int(max_tile_exp))
        if torch.cuda.current_device() == self.index:
            rc = self.lib.g2048_one(*args, _raw_stream(self.index), 1)
        else:
            with torch.cuda.device(self.index):
                rc = self.lib.g2048_one(*args, _raw_stream(self.index), 1)
        if rc:
            check(rc)
        if need_valid and self.io.bad_cells:
            raise ValueError("board holds a cell that is not 0 or a power of two: %s" % (np.asarray(matrix),))
        return self.io


def stack(flat, layers=15):
    """[4,4] tile values -> [layers+1,4,4] one-hot (reference :17-32): channel 0 = empty,
    channel k = (cell == 2**k); any other value lights no channel.  One g2048_one call."""
    flat = np.asarray(flat)
    if layers != 15 or flat.shape != (4, 4):
        raise ValueError("stack() supports the reference's 4x4 board with layers=15")
    d = _Device.get()
    d.call(_lib.ONE_STATUS, flat, need_valid=False)
    return d.obs.astype(int)


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

    def _write_back(self, d):
        # through the caller's array: set_board() aliases it (:286-288) and move() mutates it in place
        self.Matrix[...] = d.values.reshape(4, 4)

    # -- gymnasium interface -------------------------------------------------------------
    def step(self, action):
        """(:76-100) move, add a tile, detect the end.  An illegal move terminates (:91-95).
        One kernel launch: the Matrix goes in and the new Matrix, reward, flags, highest and the
        observation come back in the same pinned block."""
        d = _Device.get()
        info = {'illegal_move': False}
        io = d.call(_lib.ONE_STEP, self.Matrix, action=action, seed=self._seed, index=self._step_index,
                    illegal_move_reward=self.illegal_move_reward, max_tile_exp=tile_to_exp(self.max_tile))
        self._step_index += 1
        terminated = bool(io.done)
        if io.illegal:
            info['illegal_move'] = True
            reward = self.illegal_move_reward
        else:
            reward = float(io.reward)
            assert reward <= 2**(self.w * self.h)                                          # :87
            self.score += reward
            self._write_back(d)
        e = int(io.highest_exp)
        info['highest'] = np.int64((1 << e) if e else 0)                                   # :97
        return d.obs.astype(int), reward, terminated, False, info

    def reset(self, seed=None, options=None):
        """(:102-111) zero board, score 0, two tiles.  `seed` re-keys the draw stream."""
        if _gym is not None:
            super().reset(seed=seed)
        if seed is not None:
            self._seed = int(seed) & (2**64 - 1)
            self._step_index = 0
            self._reset_index = 0
        d = _Device.get()
        d.call(_lib.ONE_RESET, None, seed=self._seed, index=self._reset_index)
        self._reset_index += 1
        self.Matrix = np.zeros((self.h, self.w), int)
        self._write_back(d)
        self.score = 0
        return d.obs.astype(int), {}

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
        d.call(_lib.ONE_ADD_TILE, self.Matrix, seed=self._seed, index=self._step_index)
        self._step_index += 1
        self._write_back(d)

    def get(self, x, y):
        return self.Matrix[x, y]

    def set(self, x, y, val):
        self.Matrix[x, y] = val

    def empties(self):
        return np.argwhere(self.Matrix == 0)

    def _status(self):
        return _Device.get().call(_lib.ONE_STATUS, self.Matrix, max_tile_exp=tile_to_exp(self.max_tile))

    def highest(self):
        """(:190-192) the highest tile on the board (numpy int, like np.max)."""
        e = int(self._status().highest_exp)
        return np.int64((1 << e) if e else 0)

    def legal_actions(self):
        """Bit mask of legal moves (bit d <=> move(d, trial=True) does not raise)."""
        return int(self._status().legal_mask)

    def move(self, direction, trial=False):
        """(:194-241) slide/merge toward 0=up 1=right 2=down 3=left; returns the score, raises
        IllegalMove when nothing moves.  trial=True leaves the board untouched."""
        d = _Device.get()
        io = d.call(_lib.ONE_MOVE, self.Matrix, action=direction, trial=trial)
        if not io.changed:
            raise IllegalMove
        if not trial:
            self._write_back(d)
        return int(io.score)

    def shift(self, row):
        """(:243-260) compact and combine one line toward index 0: ([4 values], score)."""
        row = [int(v) for v in row]
        assert len(row) == self.size
        d = _Device.get()
        board = np.zeros(16, np.int64)
        board[:4] = row                                              # the line as row 0, moved Left
        try:
            io = d.call(_lib.ONE_MOVE, board, action=3)
        except ValueError:
            raise ValueError("shift() needs tile values that are 0 or powers of two")
        return ([int(v) for v in d.values[:4]], int(io.score))

    def isend(self):
        """(:262-280) max_tile reached, or no empty cell and no legal move."""
        return bool(self._status().is_end)

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
