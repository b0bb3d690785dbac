"""Transition data either side of the step: a device-resident recorder and the reference's
CSV schema.

Mirrors the parts of `/root/reference/training_data.py` that touch the step path's data —
`hflip` (:257-272), `rotate` (:274-279), `augment` (:281-299), `get_discounted_return`
(:104-124), `export_csv` / `import_csv` (:188-248), `merge`, `size`, the getters — over
tensors on the GPU, and the recording rule of `gather_training_data.py:191-196` (one row per
step: board, action, reward, next board, done; illegal moves are not recorded).  Boards are
16-byte exponent boards on the device and 4x4 integer tile values at the surface and in the
CSV file, so files written here load with the reference's `training_data.import_csv` and the
reverse.  All transforms run in libg2048.so (no CPU fallback).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import G2048Error, check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Transitions:
    """n rows of (board, action, reward, next_board, done) on one GPU."""

    def __init__(self, boards=None, actions=None, rewards=None, next_boards=None, dones=None, device=None):
        if not torch.cuda.is_available():
            raise G2048Error("Transitions needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.lib()
        if device is None:
            device = boards.device if isinstance(boards, torch.Tensor) and boards.is_cuda else \
                torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        n = 0 if boards is None else int(torch.as_tensor(boards).reshape(-1, 16).shape[0])

        def col(x, dtype, shape):
            if x is None:
                return torch.zeros(shape, dtype=dtype, device=self.device)
            t = torch.as_tensor(x)
            if t.dtype == torch.bool:
                t = t.to(torch.uint8)
            return t.to(device=self.device, dtype=dtype).reshape(shape).contiguous()
        self.boards = col(boards, torch.uint8, (n, 16))
        self.actions = col(actions, torch.uint8, (n,))
        self.rewards = col(rewards, torch.float32, (n,))
        self.next_boards = col(next_boards, torch.uint8, (n, 16))
        self.dones = col(dones, torch.uint8, (n,))

    # -- training_data getters (tile VALUES, the reference's shapes) ---------------------------
    def size(self):
        return int(self.boards.shape[0])

    def _values(self, b):
        out = torch.empty((b.shape[0], 4, 4), dtype=torch.int64, device=self.device)
        if b.shape[0]:
            with torch.cuda.device(self.device):
                check(self.lib.g2048_values_from_exp(_ptr(b), _ptr(out), b.numel(), self._stream()))
        return out

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def get_x(self):
        return self._values(self.boards)

    def get_next_x(self):
        return self._values(self.next_boards)

    def get_y_digit(self):
        return self.actions.to(torch.int64).reshape(-1, 1)

    def get_reward(self):
        return self.rewards.to(torch.float64).reshape(-1, 1)

    def get_done(self):
        return self.dones.to(torch.bool).reshape(-1, 1)

    def get_total_reward(self):
        return float(self.rewards.to(torch.float64).sum())

    def get_highest_tile(self):
        return int(1 << int(self.next_boards.max())) if self.size() and int(self.next_boards.max()) else 0

    def copy(self):
        return Transitions(self.boards.clone(), self.actions.clone(), self.rewards.clone(), self.next_boards.clone(),
                           self.dones.clone(), device=self.device)

    def merge(self, other):
        self.boards = torch.cat([self.boards, other.boards.to(self.device)])
        self.actions = torch.cat([self.actions, other.actions.to(self.device)])
        self.rewards = torch.cat([self.rewards, other.rewards.to(self.device)])
        self.next_boards = torch.cat([self.next_boards, other.next_boards.to(self.device)])
        self.dones = torch.cat([self.dones, other.dones.to(self.device)])

    def sample(self, index_list):
        idx = torch.as_tensor(index_list, dtype=torch.int64, device=self.device)
        return Transitions(self.boards[idx], self.actions[idx], self.rewards[idx], self.next_boards[idx],
                           self.dones[idx], device=self.device)

    # -- symmetries (:257-299) -----------------------------------------------------------------
    def _symmetry(self, hflip, k):
        n = self.size()
        if n == 0:
            return
        with torch.cuda.device(self.device):
            check(self.lib.g2048_symmetry(_ptr(self.boards), _ptr(self.boards), _ptr(self.next_boards),
                                          _ptr(self.next_boards), _ptr(self.actions), _ptr(self.actions), n,
                                          int(bool(hflip)), int(k), self._stream()))

    def hflip(self):
        """Flip all the data horizontally; directions 1 and 3 swap (:257-272)."""
        self._symmetry(True, 0)

    def rotate(self, k):
        """Rotate the boards by k quarter turns, np.rot90(axes=(2,1)); action + k mod 4 (:274-279)."""
        self._symmetry(False, int(k) % 4)

    def augment(self):
        """The 8 symmetric copies in the reference's order [X, H, R1 X, R1 H, ..., R3 H] (:281-299)."""
        n = self.size()
        if n == 0:
            return
        dev = self.device
        ob = torch.empty((8 * n, 16), dtype=torch.uint8, device=dev)
        onb = torch.empty((8 * n, 16), dtype=torch.uint8, device=dev)
        oa = torch.empty(8 * n, dtype=torch.uint8, device=dev)
        orw = torch.empty(8 * n, dtype=torch.float32, device=dev)
        od = torch.empty(8 * n, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.g2048_augment(_ptr(self.boards), _ptr(self.next_boards), _ptr(self.actions),
                                         _ptr(self.rewards), _ptr(self.dones), n, _ptr(ob), _ptr(onb), _ptr(oa),
                                         _ptr(orw), _ptr(od), self._stream()))
        self.boards, self.next_boards, self.actions, self.rewards, self.dones = ob, onb, oa, orw, od

    # -- returns (:104-124) ----------------------------------------------------------------------
    def get_discounted_return(self, gamma=0.9):
        """float64 [n,1]; relies on the rows being in game order, `done` ends an episode."""
        n = self.size()
        out = torch.zeros(n, dtype=torch.float64, device=self.device)
        if n:
            with torch.cuda.device(self.device):
                check(self.lib.g2048_discounted_return(_ptr(self.rewards), _ptr(self.dones), _ptr(out), n,
                                                       float(gamma), self._stream()))
        return out.reshape(-1, 1)

    # -- CSV (:188-248) --------------------------------------------------------------------------
    def export_csv(self, filename, add_returns=False, append=False):
        """Write the reference's 35(+1)-column CSV (tile values, '%f' reward)."""
        ret = self.get_discounted_return().reshape(-1).cpu().numpy() if add_returns else None
        b = np.ascontiguousarray(self.boards.cpu().numpy())
        a = np.ascontiguousarray(self.actions.cpu().numpy())
        r = np.ascontiguousarray(self.rewards.cpu().numpy().astype(np.float64))
        nb = np.ascontiguousarray(self.next_boards.cpu().numpy())
        d = np.ascontiguousarray(self.dones.cpu().numpy())
        p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)      # noqa: E731
        check(self.lib.g2048_csv_export(str(filename).encode(), p(b), p(a), p(r), p(nb), p(d), p(ret), len(a),
                                        1 if append else 0))

    def import_csv(self, filename):
        """Load a CSV written by this class or by the reference's training_data.export_csv."""
        path = str(filename).encode()
        n, has_ret = C.c_uint64(0), C.c_int(0)
        check(self.lib.g2048_csv_rows(path, C.byref(n), C.byref(has_ret)))
        n = int(n.value)
        b, nb = np.zeros((n, 16), np.uint8), np.zeros((n, 16), np.uint8)
        a, d = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        r = np.zeros(n, np.float64)
        p = lambda x: x.ctypes.data_as(C.c_void_p)                              # noqa: E731
        check(self.lib.g2048_csv_import(path, p(b), p(a), p(r), p(nb), p(d), None, n))
        dev = self.device
        self.boards, self.next_boards = torch.from_numpy(b).to(dev), torch.from_numpy(nb).to(dev)
        self.actions, self.dones = torch.from_numpy(a).to(dev), torch.from_numpy(d).to(dev)
        self.rewards = torch.from_numpy(r.astype(np.float32)).to(dev)
        return self


class TransitionRecorder:
    """Records every step of a BatchedGame2048 into device buffers with no copy on the step
    path: the env steps OUT OF PLACE from slice t to slice t+1 of a [T+1,n,16] board trajectory
    (G2048StepArgs.boards_out), terminal boards land in their own [T,n,16] slice."""

    def __init__(self, game, horizon):
        self.game, self.T = game, int(horizon)
        n, dev = game.num_envs, game.device
        self.traj = torch.zeros((self.T + 1, n, 16), dtype=torch.uint8, device=dev)
        self.terminal = torch.zeros((self.T, n, 16), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((self.T, n), dtype=torch.uint8, device=dev)
        self.rewards = torch.zeros((self.T, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((self.T, n), dtype=torch.uint8, device=dev)
        self.illegal = torch.zeros((self.T, n), dtype=torch.uint8, device=dev)
        self.t = 0
        if game._illegal is None:
            raise ValueError("TransitionRecorder needs the 'illegal' output of the game")
        self.traj[0].copy_(game.boards)
        game.boards = self.traj[0]

    def step(self, actions):
        """game.step(actions), recorded.  Returns the StepResult."""
        if self.t >= self.T:
            raise IndexError("recorder is full (%d steps); call transitions() / rewind()" % self.T)
        t = self.t
        res = self.game.step(actions, boards_out=self.traj[t + 1], terminal_out=self.terminal[t])
        self.actions[t].copy_(torch.as_tensor(actions).to(self.game.device).to(torch.uint8))
        self.rewards[t].copy_(res.rewards)
        self.dones[t].copy_(res.dones)
        self.illegal[t].copy_(res.illegal)
        self.t += 1
        return res

    def rewind(self):
        """Start a new recording at the current boards."""
        self.traj[0].copy_(self.game.boards)
        self.game.boards = self.traj[0]
        self.t = 0

    def transitions(self, drop_illegal=True):
        """The recorded rows in game order (env-major: all steps of env 0, then env 1, ...), as the
        reference's gather loop appends them; illegal moves are not recorded (:193-196)."""
        t = self.t
        done = self.dones[:t].to(torch.bool)
        nxt = torch.where(done[..., None], self.terminal[:t], self.traj[1:t + 1])
        em = lambda x: x.transpose(0, 1).reshape((-1,) + tuple(x.shape[2:]))      # noqa: E731
        b, a, r, nb, d = em(self.traj[:t]), em(self.actions[:t]), em(self.rewards[:t]), em(nxt), em(self.dones[:t])
        if drop_illegal:
            keep = em(self.illegal[:t]) == 0
            b, a, r, nb, d = b[keep], a[keep], r[keep], nb[keep], d[keep]
        return Transitions(b.contiguous(), a.contiguous(), r.contiguous(), nb.contiguous(), d.contiguous(),
                           device=self.game.device)
