"""BatchedGame2048 — N independent 2048 boards stepped by one CUDA kernel per call.

Host-side mirror of the reference env's interface for the batched path
(`/root/reference/env/envs/game2048_env.py`: reset :102-111, step :76-100, move :194-241,
isend :262-280, highest :190-192, stack :17-32, set_illegal_move_reward :61-67,
set_max_tile :69-73), with SB3 `DummyVecEnv` same-step auto-reset.  torch is used only for
device memory and streams; every game rule runs in libg2048.so (no CPU fallback).
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import StepManyArgs, FLAG_AUTO_RESET, G2048Error, StepArgs, check

try:                                    # ~0.1 us instead of ~2 us for torch.cuda.current_stream().cuda_stream
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                  # pragma: no cover - older/newer torch without the private hook
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream

ALL_OUTPUTS = ("illegal", "highest", "legal_mask", "episode", "terminal")
_OBS_DTYPES = {torch.uint8: _lib.OBS_U8, torch.float32: _lib.OBS_F32, torch.int64: _lib.OBS_I64,
               torch.bfloat16: _lib.OBS_BF16}


def tile_to_exp(max_tile):
    """set_max_tile value -> exponent for the kernel (0 = None).  A value that no tile can
    ever equal (not a power of two >= 2) maps to 63, which never matches (`==`, :267)."""
    if max_tile is None:
        return 0
    assert isinstance(max_tile, int), "max_tile must be an int or None"      # :72
    if max_tile >= 2 and (max_tile & (max_tile - 1)) == 0 and max_tile.bit_length() - 1 < 63:
        return max_tile.bit_length() - 1
    return 63


def shard_range(total, rank, world):
    """Contiguous slice [base, base+count) of `total` global env ids owned by `rank`."""
    per, rem = divmod(int(total), int(world))
    base = rank * per + min(rank, rem)
    return base, per + (1 if rank < rem else 0)


@dataclass
class StepResult:
    boards: torch.Tensor                      # uint8 [n,16] exponents — the env's live state
    rewards: torch.Tensor                     # float32 [n]
    dones: torch.Tensor                       # bool [n]
    illegal: Optional[torch.Tensor] = None    # bool [n]   info['illegal_move']
    highest_exp: Optional[torch.Tensor] = None  # uint8 [n] log2(info['highest'])
    legal_mask: Optional[torch.Tensor] = None   # uint8 [n] bit d = move d legal on `boards`
    terminal_boards: Optional[torch.Tensor] = None  # uint8 [n,16], valid where dones
    final_score: Optional[torch.Tensor] = None  # int32 [n] episode score, valid where dones
    final_len: Optional[torch.Tensor] = None    # int32 [n] episode length, valid where dones
    final_return: Optional[torch.Tensor] = None  # float32 [n] sum of the episode's rewards (SB3 Monitor 'r'), where dones
    actions: Optional[torch.Tensor] = None      # uint8 [n] the actions played (step(policy=...): drawn by the kernel)


class BatchedGame2048:
    """`num_envs` boards resident on one GPU.

    env_id_base is the global id of board 0: draws depend on (seed, global id, step), so a
    batch sharded over several GPUs/processes reproduces the unsharded result bit for bit.
    """

    def __init__(self, num_envs, seed=0, device=None, env_id_base=0, illegal_move_reward=0.0,
                 max_tile=None, auto_reset=True, outputs=ALL_OUTPUTS):
        if not torch.cuda.is_available():
            raise G2048Error("BatchedGame2048 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.env_id_base = int(env_id_base)
        self.seed = int(seed) & (2**64 - 1)
        self.auto_reset = bool(auto_reset)
        self.step_index = 0
        self.reset_index = 0
        self.set_illegal_move_reward(illegal_move_reward)
        self.set_max_tile(max_tile)
        unknown = set(outputs) - set(ALL_OUTPUTS)
        if unknown:
            raise ValueError("unknown outputs %s" % sorted(unknown))
        self.outputs = tuple(outputs)
        n, dev = self.num_envs, self.device
        u8 = dict(dtype=torch.uint8, device=dev)
        self.boards = torch.zeros((n, 16), **u8)
        self.rewards = torch.zeros(n, dtype=torch.float32, device=dev)
        self._dones = torch.zeros(n, **u8)
        self._illegal = torch.zeros(n, **u8) if "illegal" in outputs else None
        self.highest_exp = torch.zeros(n, **u8) if "highest" in outputs else None
        self.legal_mask = torch.zeros(n, **u8) if "legal_mask" in outputs else None
        self.terminal_boards = torch.zeros((n, 16), **u8) if "terminal" in outputs else None
        ep = "episode" in outputs
        i32 = dict(dtype=torch.int32, device=dev)
        self.ep_score = torch.zeros(n, **i32) if ep else None
        self.ep_len = torch.zeros(n, **i32) if ep else None
        self.final_score = torch.zeros(n, **i32) if ep else None
        self.final_len = torch.zeros(n, **i32) if ep else None
        self.ep_return = torch.zeros(n, dtype=torch.float32, device=dev) if ep else None
        self.final_return = torch.zeros(n, dtype=torch.float32, device=dev) if ep else None
        self._step_counter = None      # device uint64 step index (use_device_step_counter)
        self._actions_out = None       # where step(policy=...) writes the actions it drew
        self._chain = None             # chain buffer of chained launches (step(chained=True), step_n, StepSchedule)
        self._chain_broken = True      # the next step must not chain: something else wrote the env's state
        self._chain_mode = "direct"    # launch shape of the env's chain: "direct" or "interleaved"
        self._args = None              # cached G2048StepArgs, rebuilt when a knob changes

    # -- knobs (reference :61-73) ------------------------------------------------------
    def set_illegal_move_reward(self, reward):
        self.illegal_move_reward = float(reward)
        self.reward_range = (self.illegal_move_reward, float(2 ** 16))

    def set_max_tile(self, max_tile):
        self.max_tile = max_tile
        self.max_tile_exp = tile_to_exp(max_tile)

    # -- helpers -----------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_raw_stream(self.device.index))

    def _launch(self, fn, *args):
        """fn(*args, current stream of self.device) with that device current; raises on error."""
        idx = self.device.index
        if torch.cuda.current_device() == idx:
            rc = fn(*args, _raw_stream(idx))
        else:
            with torch.cuda.device(idx):
                rc = fn(*args, _raw_stream(idx))
        if rc:
            check(rc)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _as_u8(self, x, shape, what):
        t = torch.as_tensor(x)
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        if t.dtype != torch.uint8:
            t = t.to(torch.int64)
            if what == "actions" and t.numel() and (int(t.min()) < 0 or int(t.max()) > 3):
                raise ValueError("actions must be in {0,1,2,3} (Discrete(4))")
            t = t.to(torch.uint8)
        t = t.to(self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (what, tuple(shape), tuple(t.shape)))
        return t

    # -- reset (:102-111) --------------------------------------------------------------
    def reset(self, seed=None, mask=None):
        """Fresh boards (two spawned tiles) for all envs, or those where `mask` is set.
        A `seed` re-keys the draw stream and restarts its counters (reset(seed), :103)."""
        if seed is not None:
            self.seed = int(seed) & (2**64 - 1)
            self.step_index = 0
            self.reset_index = 0
            if self._step_counter is not None:
                self._step_counter.zero_()
        m = None if mask is None else self._as_u8(mask, (self.num_envs,), "mask")
        self._chain_broken = True
        self._launch(self.lib.g2048_reset, self._ptr(self.boards), self._ptr(m), self.num_envs, self.env_id_base,
                     self.seed, self.reset_index)
        self.reset_index += 1
        if self.ep_score is not None:
            if m is None:
                self.ep_score.zero_()
                self.ep_len.zero_()
                self.ep_return.zero_()
            else:
                keep = (m == 0).to(torch.int32)
                self.ep_score.mul_(keep)
                self.ep_len.mul_(keep)
                self.ep_return.mul_(keep.to(torch.float32))
        if self.legal_mask is not None:
            self.status(legal_mask=self.legal_mask)
        return self.boards

    # -- step (:76-100) ----------------------------------------------------------------
    def _build_step_args(self):
        a = StepArgs(self._ptr(self.boards), None, self._ptr(self.rewards), self._ptr(self._dones),
                     self._ptr(self._illegal), self._ptr(self.highest_exp), self._ptr(self.legal_mask),
                     self._ptr(self.terminal_boards), self._ptr(self.ep_score), self._ptr(self.ep_len),
                     self._ptr(self.final_score), self._ptr(self.final_len), None, self._ptr(self._step_counter),
                     self.num_envs, self.env_id_base, self.seed, self.step_index,
                     self.illegal_move_reward, self.max_tile_exp, FLAG_AUTO_RESET if self.auto_reset else 0, None,
                     self._ptr(self.ep_return), self._ptr(self.final_return))
        self._args = a
        self._args_ref = C.byref(a)
        self._args_key = (self.seed, self.env_id_base, self.illegal_move_reward, self.max_tile_exp,
                          self.auto_reset, None if self._step_counter is None else self._step_counter.data_ptr())
        self._result = StepResult(self.boards, self.rewards, self._dones.view(torch.bool),
                                  None if self._illegal is None else self._illegal.view(torch.bool),
                                  self.highest_exp, self.legal_mask, self.terminal_boards, self.final_score,
                                  self.final_len, self.final_return)

    def _chain_args(self, a, chained):
        """Chained launches (G2048StepArgs.chain / G2048_FLAG_CHAINED) for the step call `a`; chained = False, True or
        "interleaved" (see step()).  The env's chain buffer is made (zeroed) at the first chained step.  A chained step
        that follows anything else this class did to the env's state (reset, set_boards, step_many, a plain step,
        ...), or that changes between direct and interleaved chaining (another launch shape), carries the buffer but
        not the flag: it waits for all earlier work like any launch and publishes, so the NEXT step can chain to it.
        Plain steps are launched without the buffer, in the full-machine shape."""
        if chained not in (False, True, "interleaved"):
            raise ValueError("chained must be False, True or 'interleaved'")
        if not chained or self._step_counter is not None:
            if chained:
                raise G2048Error("chained steps are not available with a device-side step counter")
            # a plain launch: the full-machine shape, outside the protocol — the next chained step starts a new chain
            a.chain = None
            self._chain_broken = True
            return
        if self._chain is None:
            self._chain = torch.zeros(_lib.CHAIN_WORDS, dtype=torch.int64, device=self.device)
            self._chain_broken = True
        a.chain = self._chain.data_ptr()
        mode = "interleaved" if chained == "interleaved" else "direct"
        if mode == "interleaved":
            a.flags |= _lib.FLAG_CHAIN_INTERLEAVED
        if not self._chain_broken and mode == self._chain_mode:
            a.flags |= _lib.FLAG_CHAINED
        self._chain_broken = False
        self._chain_mode = mode

    def _check_board_buffer(self, t, what):
        if not (isinstance(t, torch.Tensor) and t.dtype == torch.uint8 and t.device == self.device
                and t.is_contiguous() and tuple(t.shape) == (self.num_envs, 16)):
            raise ValueError("%s must be a contiguous uint8 [%d,16] tensor on %s" % (what, self.num_envs, self.device))
        return t

    def step(self, actions=None, forced_draws=None, boards_out=None, terminal_out=None, policy=None, chained=False):
        """One env step for every board.  `actions`: uint8/int tensor [n] on any device.

        chained=True / "interleaved" (G2048_FLAG_CHAINED, include/g2048.h): the caller's promise that everything this
        step reads — the boards, `actions` (or the legal mask with policy='legal'), the episode statistics — was
        last written by THIS env's previous step() or before that step was issued: open-loop action rows prepared
        in advance, the in-kernel policies, several envs stepped round-robin.  The launch then depends on its
        predecessor warp by warp instead of waiting for the whole grid to drain, with identical results.  True: this
        env is stepped back to back (1 Mi boards 10.5 -> 9.4 us per step, 262,144 4.1 -> 3.6 us); "interleaved":
        steps of OTHER envs sit between two steps of this one (two or more env sets round-robin: 10.8 -> 9.05 us and
        4.4 -> 2.5 us; with a single set this shape is slower than plain launches, profiles/r02_chain_sets.log).  Leave it False when a policy kernel wrote `actions` after the previous step (a closed
        loop), or after touching the env's tensors directly.

        policy = "uniform" / "legal": the kernel draws the actions itself — exactly the actions sample_actions(legal=...)
        would return for this step — and steps with them in the SAME launch (BASELINE config 4: sample a legal action,
        step, hand back the new legal mask, one kernel per step); they are written to `actions` when given (a
        contiguous uint8 [n] tensor on the device) or to an internal buffer, and come back as StepResult.actions.

        boards_out: step OUT OF PLACE — the boards handed back to the agent are written there
        and become the env's live state (`self.boards`), the previous buffer keeps the pre-step
        boards.  A trajectory buffer [T+1,n,16] is filled by passing slice t+1 at step t, with no
        copy.  terminal_out: where this step's terminal boards go instead of `terminal_boards`.

        The returned StepResult holds the env's own output tensors (overwritten by the next
        step); clone what must outlive it."""
        n = self.num_envs
        pol = 0
        if policy is not None:
            if policy not in ("uniform", "legal"):
                raise ValueError("policy must be None, 'uniform' or 'legal'")
            if policy == "legal" and self.legal_mask is None:
                raise ValueError("step(policy='legal') needs the 'legal_mask' output")
            if forced_draws is not None:
                raise ValueError("forced_draws and a policy cannot be combined")
            pol = _lib.FLAG_POLICY_LEGAL if policy == "legal" else _lib.FLAG_POLICY_UNIFORM
            if actions is None:
                if self._actions_out is None:
                    self._actions_out = torch.empty(n, dtype=torch.uint8, device=self.device)
                act = self._actions_out
            elif isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8 and actions.device == self.device \
                    and actions.is_contiguous() and actions.shape == (n,):
                act = actions
            else:
                raise ValueError("with a policy, actions is an OUTPUT: a contiguous uint8 [%d] tensor on %s" % (n, self.device))
        elif isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8 and actions.device == self.device \
                and actions.is_contiguous() and actions.shape == (n,):
            act = actions                       # fast path: no validation pass over the batch
        else:
            if actions is None:
                raise ValueError("step() needs actions (or policy='uniform' / 'legal')")
            act = self._as_u8(actions, (n,), "actions")
        key = (self.seed, self.env_id_base, self.illegal_move_reward, self.max_tile_exp, self.auto_reset,
               None if self._step_counter is None else self._step_counter.data_ptr())
        if self._args is None or key != self._args_key:
            self._build_step_args()
        a = self._args
        a.actions = act.data_ptr()
        a.step_index = self.step_index
        a.boards = self.boards.data_ptr()
        a.flags = (FLAG_AUTO_RESET if self.auto_reset else 0) | pol
        self._chain_args(a, chained)
        a.boards_out = None if boards_out is None else self._check_board_buffer(boards_out, "boards_out").data_ptr()
        # every per-call pointer of the cached struct is reassigned on every call: a stale terminal_out would
        # keep the kernel writing into a buffer its owner (e.g. a dropped TransitionRecorder) may have freed
        if terminal_out is not None:
            a.terminal_boards = self._check_board_buffer(terminal_out, "terminal_out").data_ptr()
        else:
            a.terminal_boards = None if self.terminal_boards is None else self.terminal_boards.data_ptr()
        fd = None
        if forced_draws is not None:
            fd = torch.as_tensor(forced_draws).to(self.device).contiguous()
            if fd.dtype != torch.uint32 or tuple(fd.shape) != (n, 4):
                raise ValueError("forced_draws must be uint32 [n,4]")
            a.forced_draws = fd.data_ptr()
        else:
            a.forced_draws = None
        self._launch(self.lib.g2048_step, self._args_ref)
        self.step_index += 1      # host mirror; the device counter (if any) is bumped on the stream
        res = self._result
        if boards_out is not None:
            self.boards = boards_out
        res.boards = self.boards
        res.terminal_boards = terminal_out if terminal_out is not None else self.terminal_boards
        res.actions = act
        return res

    def step_n(self, actions, rewards=None, dones=None, illegal=None, highest_exp=None, legal_mask_out=None,
               chained=True):
        """K consecutive steps, ONE KERNEL LAUNCH PER STEP, issued back to back from C (g2048_step_n): what
        `for k in range(K): step(actions[k])` does — every step reads and writes the boards in device memory —
        without the interpreter between two launches (a Python loop issues a launch every ~6 us, a B200 steps
        131,072 boards in ~3 us).  `actions`: contiguous uint8 [K,n] on the device (an open-loop sequence).
        Returns (rewards f32 [K,n], dones bool [K,n]); `illegal`, `highest_exp`, `legal_mask_out` (uint8 [K,n])
        are filled when given.  The env's running episode statistics (ep_score/ep_len/ep_return) are kept;
        terminal boards and final_* of the intermediate steps are not reported (use step() for those).
        chained (default): steps 2..K are chained launches (see step()) — every action row was there before the
        call, so step k+1 only depends on step k, slice by slice; the first step waits for all earlier work on the
        stream as any launch does."""
        n = self.num_envs
        if self._step_counter is not None:
            raise G2048Error("step_n is not available with a device-side step counter")
        if not (isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8 and actions.device == self.device
                and actions.is_contiguous() and actions.dim() == 2 and actions.shape[1] == n):
            raise ValueError("actions must be a contiguous uint8 [K,%d] tensor on %s" % (n, self.device))
        K = int(actions.shape[0])

        def buf(t, dtype, what, alloc):
            if t is None:
                return torch.empty((K, n), dtype=dtype, device=self.device) if alloc else None
            if not (isinstance(t, torch.Tensor) and t.dtype == dtype and t.device == self.device and t.is_contiguous()
                    and tuple(t.shape) == (K, n)):
                raise ValueError("%s must be a contiguous %s [%d,%d] tensor on %s" % (what, dtype, K, n, self.device))
            return t
        rewards = buf(rewards, torch.float32, "rewards", True)
        dones_u8 = buf(dones, torch.uint8, "dones", True)
        illegal = buf(illegal, torch.uint8, "illegal", False)
        highest_exp = buf(highest_exp, torch.uint8, "highest_exp", False)
        if legal_mask_out is None and self.legal_mask is not None and K:
            legal_mask_out = torch.empty((K, n), dtype=torch.uint8, device=self.device)
        legal_mask_out = buf(legal_mask_out, torch.uint8, "legal_mask_out", False)
        if K == 0:
            return rewards, dones_u8.view(torch.bool)
        a = StepArgs(self._ptr(self.boards), self._ptr(actions), self._ptr(rewards), self._ptr(dones_u8),
                     self._ptr(illegal), self._ptr(highest_exp), self._ptr(legal_mask_out), None,
                     self._ptr(self.ep_score), self._ptr(self.ep_len), None, None, None, None,
                     n, self.env_id_base, self.seed, self.step_index,
                     self.illegal_move_reward, self.max_tile_exp, FLAG_AUTO_RESET if self.auto_reset else 0, None,
                     self._ptr(self.ep_return), None)
        if chained:
            self._chain_args(a, True)          # (the C loop sets the flag itself from the second step on)
        self._launch(self.lib.g2048_step_n, C.byref(a), K, n)
        self.step_index += K
        if self.legal_mask is not None:
            self.legal_mask.copy_(legal_mask_out[K - 1])
        self.rewards.copy_(rewards[K - 1])
        self._dones.copy_(dones_u8[K - 1])
        return rewards, dones_u8.view(torch.bool)

    def capture(self, fn, warmup=1):
        """Capture `fn()` — any fixed sequence of step() / observe() / sample_actions-free policy work on this
        env's device, e.g. `policy forward -> step` repeated T times — in a CUDA graph and return a callable
        that replays it: one graph launch instead of one trip through the interpreter per kernel.  This is the
        closed-loop answer to small batches (a 65,536-board step takes ~3 us on the GPU and ~4-6 us to issue
        from Python); for open-loop action sequences step_n / step_many need no graph.

        The env switches to its DEVICE-side step index (use_device_step_counter): the captured step kernels read
        the index from device memory and advance it themselves, so every replay draws fresh tiles.  `fn` must
        only touch tensors that stay alive and in place (write new actions INTO the tensors fn read);
        `warmup` eager calls of fn run first (they are real steps).  Returns replay(); replay.graph is the
        torch.cuda.CUDAGraph, replay.steps the number of env steps per replay."""
        if self._step_counter is None:
            self.use_device_step_counter(True)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(int(warmup)):
                fn()
            before = self.step_index
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                fn()
            steps = self.step_index - before
            self.step_index = before            # capture does not execute
        cur.wait_stream(side)

        def replay():
            graph.replay()
            self.step_index += steps
        replay.graph, replay.steps = graph, steps
        return replay

    def step_many(self, actions=None, rewards=None, dones=None, illegal=None, boards_traj=None, policy=None,
                  n_steps=None, actions_out=None, legal_mask_out=None):
        """K steps in ONE launch (g2048_step_many), bit-identical to K calls of step() — same draws, same
        auto-reset — with the boards staying in registers between the steps.

        Either `actions` (uint8 [K,n] on the device: an open-loop sequence — pre-generated random actions, a
        recorded game), or `policy` = "uniform" / "legal" with `n_steps` = K: the kernel then plays
        sample_actions()'s policy itself (the action of step k is what sample_actions(legal=...) would return
        before that step) and writes the actions it chose to `actions_out` (uint8 [K,n]) when given.
        Returns (rewards f32 [K,n], dones bool [K,n]); `illegal` (uint8 [K,n]), `boards_traj` (uint8 [K,n,16],
        the board handed back after every step) and `legal_mask_out` (uint8 [K,n], its legal moves) are filled
        when given.  step()'s other optional outputs (highest, episode statistics) are not produced: the env's
        legal mask is refreshed at the end, episode statistics must not be enabled."""
        n = self.num_envs
        if self._step_counter is not None:
            raise G2048Error("step_many is not available with a device-side step counter")
        if self.ep_score is not None:
            raise G2048Error("step_many does not maintain episode statistics: build the env without 'episode'")
        flags = FLAG_AUTO_RESET if self.auto_reset else 0
        if policy is None:
            if not (isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8 and actions.device == self.device
                    and actions.is_contiguous() and actions.dim() == 2 and actions.shape[1] == n):
                raise ValueError("actions must be a contiguous uint8 [K,%d] tensor on %s" % (n, self.device))
            K = int(actions.shape[0])
            if n_steps is not None and int(n_steps) != K:
                raise ValueError("n_steps disagrees with actions.shape[0]")
            if actions_out is not None:
                raise ValueError("actions_out is only written when a policy draws the actions")
        else:
            if policy not in ("uniform", "legal"):
                raise ValueError("policy must be None, 'uniform' or 'legal'")
            if actions is not None or n_steps is None:
                raise ValueError("with a policy pass n_steps and no actions")
            K = int(n_steps)
            flags |= _lib.FLAG_POLICY_LEGAL if policy == "legal" else _lib.FLAG_POLICY_UNIFORM

        def buf(t, dtype, shape, what):
            if t is None:
                return None
            if not (isinstance(t, torch.Tensor) and t.dtype == dtype and t.device == self.device and t.is_contiguous()
                    and tuple(t.shape) == shape):
                raise ValueError("%s must be a contiguous %s %s tensor on %s" % (what, dtype, shape, self.device))
            return t
        rewards = buf(rewards, torch.float32, (K, n), "rewards")
        dones_u8 = buf(dones, torch.uint8, (K, n), "dones")
        if rewards is None:
            rewards = torch.empty((K, n), dtype=torch.float32, device=self.device)
        if dones_u8 is None:
            dones_u8 = torch.empty((K, n), dtype=torch.uint8, device=self.device)
        illegal = buf(illegal, torch.uint8, (K, n), "illegal")
        boards_traj = buf(boards_traj, torch.uint8, (K, n, 16), "boards_traj")
        actions_out = buf(actions_out, torch.uint8, (K, n), "actions_out")
        legal_mask_out = buf(legal_mask_out, torch.uint8, (K, n), "legal_mask_out")
        a = StepManyArgs(self._ptr(self.boards), self._ptr(actions), self._ptr(rewards), self._ptr(dones_u8),
                         self._ptr(illegal), self._ptr(boards_traj), n, self.env_id_base, self.seed, self.step_index,
                         K, self.illegal_move_reward, self.max_tile_exp, flags, self._ptr(actions_out),
                         self._ptr(legal_mask_out))
        if K:
            self._chain_broken = True
            self._launch(self.lib.g2048_step_many, C.byref(a))
            self.step_index += K
            if self.legal_mask is not None:
                if legal_mask_out is not None:
                    self.legal_mask.copy_(legal_mask_out[K - 1])
                else:
                    self.status(legal_mask=self.legal_mask)
        return rewards, dones_u8.view(torch.bool)

    def sample_actions(self, legal=False, out=None):
        """Uniform-random actions for the NEXT step, drawn on the device from the policy-tag draw stream at that
        step's index (reference: `random.randint(0, 3)`, train.py:119).  legal=True draws
        uniformly among the legal moves of the live boards (needs the `legal_mask` output)."""
        if legal and self.legal_mask is None:
            raise ValueError("sample_actions(legal=True) needs the 'legal_mask' output")
        if self._step_counter is not None:
            raise G2048Error("sample_actions is not available with a device-side step counter")
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        self._launch(self.lib.g2048_sample_actions, self._ptr(self.legal_mask) if legal else None, self._ptr(out),
                     self.num_envs, self.env_id_base, self.seed, self.step_index)
        return out

    def use_device_step_counter(self, enable=True):
        """Keep the step index in device memory (advanced by the step kernel itself as it ends) so
        a loop of step() calls can be captured once in a CUDA graph and replayed: with the
        default host-side index every replay would reuse the captured step's draws."""
        if enable:
            self._step_counter = torch.tensor([self.step_index, 0], dtype=torch.int64, device=self.device)
        else:
            if self._step_counter is not None:
                self.step_index = int(self._step_counter[0].item())
            self._step_counter = None

    # -- move / status -----------------------------------------------------------------
    def move(self, directions, trial=False):
        """Game2048Env.move for every board (no spawn).  Returns (scores int32 [n], changed bool [n]);
        changed == False is the reference's IllegalMove.  trial=True leaves the boards untouched."""
        d = self._as_u8(directions, (self.num_envs,), "actions")
        scores = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        changed = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        self._chain_broken = self._chain_broken or not trial
        self._launch(self.lib.g2048_move, self._ptr(self.boards), None if trial else self._ptr(self.boards),
                     self._ptr(d), self._ptr(scores), self._ptr(changed), self.num_envs)
        return scores, changed.view(torch.bool)

    def status(self, legal_mask=None):
        """legal-move mask, highest exponent, number of empties and isend() of the current boards."""
        n = self.num_envs
        mk = lambda: torch.zeros(n, dtype=torch.uint8, device=self.device)     # noqa: E731
        lm = mk() if legal_mask is None else legal_mask
        hi, ne, end = mk(), mk(), mk()
        self._launch(self.lib.g2048_status, self._ptr(self.boards), self._ptr(lm), self._ptr(hi), self._ptr(ne),
                     self._ptr(end), self.max_tile_exp, n)
        return dict(legal_mask=lm, highest_exp=hi, n_empty=ne, is_end=end.view(torch.bool))

    # -- observations (stack, :17-32) ----------------------------------------------------
    def observe(self, dtype=torch.uint8, out=None, boards=None):
        """One-hot [n,16,4,4] of `boards` (default: the live boards) in the reference's
        channel convention; dtype uint8 / float32 / bfloat16 / int64 (the reference's)."""
        if dtype not in _OBS_DTYPES:
            raise ValueError("unsupported obs dtype %s" % dtype)
        b = self.boards if boards is None else boards
        n = b.shape[0]
        if out is None:
            out = torch.empty((n, 16, 4, 4), dtype=dtype, device=self.device)
        elif out.dtype != dtype or tuple(out.shape) != (n, 16, 4, 4) or not out.is_contiguous():
            raise ValueError("out must be a contiguous %s tensor of shape %s" % (dtype, (n, 16, 4, 4)))
        self._launch(self.lib.g2048_encode_obs, self._ptr(b), self._ptr(out), _OBS_DTYPES[dtype], n)
        return out

    # -- tile values <-> exponents -------------------------------------------------------
    def board_values(self, boards=None):
        """int64 [n,4,4] tile values (the reference's Matrix)."""
        b = self.boards if boards is None else boards
        out = torch.empty((b.shape[0], 4, 4), dtype=torch.int64, device=self.device)
        self._launch(self.lib.g2048_values_from_exp, self._ptr(b), self._ptr(out), b.numel())
        return out

    def set_board_values(self, values):
        """Load boards given as tile values (0, 2, 4, ...), shape [n,4,4] or [n,16]."""
        v = torch.as_tensor(values).to(torch.int64).to(self.device).contiguous()
        if v.numel() != self.num_envs * 16:
            raise ValueError("expected %d cells, got %d" % (self.num_envs * 16, v.numel()))
        bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._chain_broken = True
        self._launch(self.lib.g2048_exp_from_values, self._ptr(v), self._ptr(self.boards), v.numel(), self._ptr(bad))
        if int(bad.item()):
            raise ValueError("%d cells are not 0 or a power of two in 2..2^31" % int(bad.item()))
        if self.legal_mask is not None:
            self.status(legal_mask=self.legal_mask)

    def set_boards(self, exps):
        self._chain_broken = True
        self.boards.copy_(self._as_u8(exps, (self.num_envs, 16), "boards"))
        if self.legal_mask is not None:
            self.status(legal_mask=self.legal_mask)

    # -- checkpoint / resume -------------------------------------------------------------
    def state_dict(self):
        if self._step_counter is not None:
            self.step_index = int(self._step_counter[0].item())
        sd = dict(boards=self.boards.clone(), seed=self.seed, step_index=self.step_index,
                  reset_index=self.reset_index, env_id_base=self.env_id_base)
        if self.ep_score is not None:
            sd.update(ep_score=self.ep_score.clone(), ep_len=self.ep_len.clone(), ep_return=self.ep_return.clone())
        return sd

    def load_state_dict(self, sd):
        self._chain_broken = True
        self.boards.copy_(sd["boards"])
        self.seed, self.step_index, self.reset_index = int(sd["seed"]), int(sd["step_index"]), int(sd["reset_index"])
        self.env_id_base = int(sd["env_id_base"])
        if self._step_counter is not None:
            self._step_counter[0] = self.step_index
        if self.ep_score is not None and "ep_score" in sd:
            self.ep_score.copy_(sd["ep_score"])
            self.ep_len.copy_(sd["ep_len"])
            if "ep_return" in sd:
                self.ep_return.copy_(sd["ep_return"])
        if self.legal_mask is not None:
            self.status(legal_mask=self.legal_mask)


class HostBuffers:
    """Pinned host arrays for HostSteppedEnv (numpy views over torch pinned memory)."""

    def __init__(self, n, extras=False, nibble=False):
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()       # noqa: E731
        self.actions = pin((n,), torch.uint8)
        self.boards = pin((n, 8 if nibble else 16), torch.uint8)
        self.rewards = pin((n,), torch.float32)
        self.dones = pin((n,), torch.uint8)
        self.illegal = pin((n,), torch.uint8) if extras else None
        self.highest_exp = pin((n,), torch.uint8) if extras else None
        self.legal_mask = pin((n,), torch.uint8) if extras else None
        self.nibble = nibble
        self.nibble_overflow = C.c_uint32(0)       # boards of the last step that did not fit 4 bits per cell

    def unpacked_boards(self):
        """uint8 [n,16] exponents from the compact format (cell c = nibble c of the 8 bytes), as a new tensor."""
        if not self.nibble:
            return self.boards
        out = torch.empty((self.boards.shape[0], 16), dtype=torch.uint8)
        check(_lib.lib().g2048_unpack_boards_host(C.c_void_p(self.boards.data_ptr()), C.c_void_p(out.data_ptr()),
                                                  self.boards.shape[0]))
        return out


class HostSteppedEnv:
    """The stateful host-buffer handle of the C ABI (g2048_env_*): actions come from HOST
    memory and boards/rewards/dones land in HOST memory every step, with the host<->device
    copies chunk-pipelined against the kernel inside the library.  This is the call a
    CPU-side caller (the reference's numpy world) makes; bench.py times it as `e2e`.

    board_format="nibble": the boards come back packed 4 bits per cell ([n,8] bytes, the classic 64-bit 2048
    bitboard; cell c = nibble c) — the call is PCIe-bound and this cuts the bytes per board from 21 to 13.  A board
    holding a tile >= 65,536 does not fit: `buffers.nibble_overflow.value` counts them per step and
    full_boards() fetches the 16-byte boards."""

    def __init__(self, num_envs, seed=0, device=0, env_id_base=0, illegal_move_reward=0.0, max_tile=None,
                 auto_reset=True, n_chunks=0, extras=False, board_format="bytes", wire="auto", unpack_threads=0):
        if board_format not in ("bytes", "nibble"):
            raise ValueError("board_format must be 'bytes' or 'nibble'")
        if wire not in ("auto", "packed", "plain"):
            raise ValueError("wire must be 'auto', 'packed' or 'plain'")
        if wire == "auto":         # the expansion needs host threads: with fewer than 4 CPUs the plain wire is as fast
            cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            wire = "packed" if cpus >= 4 else "plain"
        self.lib = _lib.lib()
        self.num_envs = int(num_envs)
        nibble = board_format == "nibble"
        # board_format='bytes', wire='packed' (the default on a host with 4+ CPUs): the caller still gets [n,16] exponent bytes, but they cross
        # PCIe 4 bits per cell and `unpack_threads` host threads of the library (0 = half the CPUs of the process, at
        # most 8) expand them slice by slice (G2048_BOARDS_BYTES_PACKED_WIRE); wire='plain' moves the 16 bytes.
        self.wire = "packed" if (wire == "packed" and not nibble) else "plain"
        fmt = _lib.BOARDS_NIBBLE if nibble else (_lib.BOARDS_BYTES_PACKED_WIRE if self.wire == "packed" else _lib.BOARDS_BYTES)
        cfg = _lib.EnvConfig(int(device), FLAG_AUTO_RESET if auto_reset else 0, self.num_envs, int(env_id_base),
                             int(seed) & (2**64 - 1), float(illegal_move_reward), tile_to_exp(max_tile),
                             int(n_chunks), fmt, int(unpack_threads))
        self._n_chunks = min(int(n_chunks) if n_chunks else (4 if self.wire == "packed" else 2), 64)
        self._h = C.c_void_p()
        check(self.lib.g2048_env_create(C.byref(self._h), C.byref(cfg)))
        self.buffers = HostBuffers(self.num_envs, extras, nibble)
        b = self.buffers
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())          # noqa: E731
        self._out = _lib.HostStepOut(p(b.boards), p(b.rewards), p(b.dones), p(b.illegal), p(b.highest_exp),
                                     p(b.legal_mask), C.cast(C.pointer(b.nibble_overflow), C.c_void_p))
        self._full = None

    def full_boards(self):
        """The live boards as uint8 [n,16] exponents, whatever the board format (one synchronous D2H copy)."""
        if self._full is None:
            self._full = torch.empty((self.num_envs, 16), dtype=torch.uint8).pin_memory()
        check(self.lib.g2048_env_get_boards_host(self._h, C.c_void_p(self._full.data_ptr())))
        return self._full

    def reset(self):
        """Fresh boards for every env; returns them as uint8 [n,16] exponents (also with board_format='nibble':
        the compact format applies to step results)."""
        if self.buffers.nibble:
            check(self.lib.g2048_env_reset_host(self._h, None))
            return self.full_boards()
        check(self.lib.g2048_env_reset_host(self._h, C.c_void_p(self.buffers.boards.data_ptr())))
        return self.buffers.boards

    def step(self, actions=None):
        """actions: None (already written into buffers.actions) or a uint8 array/tensor [n]."""
        if actions is not None:
            self.buffers.actions.copy_(torch.as_tensor(actions, dtype=torch.uint8))
        check(self.lib.g2048_env_step_host(self._h, C.c_void_p(self.buffers.actions.data_ptr()),
                                           C.byref(self._out)))
        return self.buffers

    def step_pinned(self, actions):
        """Like step(), reading the actions straight from the caller's (ideally pinned) host
        tensor instead of staging them through buffers.actions."""
        if actions.dtype != torch.uint8 or actions.numel() != self.num_envs or actions.is_cuda \
                or not actions.is_contiguous():
            raise ValueError("actions must be a contiguous host uint8 tensor with one entry per env")
        check(self.lib.g2048_env_step_host(self._h, C.c_void_p(actions.data_ptr()), C.byref(self._out)))
        return self.buffers

    @property
    def n_chunks_effective(self):
        """Kernel launches per step_host call: the library's slice schedule (g2048_env_step_host) — a lead
        slice of 1/16 of the boards when there are at least two slices and 65,536 boards, the rest in equal
        slices, every slice rounded up to 256 boards."""
        n, chunks = self.num_envs, self._n_chunks
        lead = -(-(n // 16) // 256) * 256 if (chunks >= 2 and n >= 65536) else 0
        tail = 0                                             # (the library's plain-tail experiment is off)
        body = n - tail
        rest = chunks - (1 if lead else 0) - (1 if tail else 0)
        per = -(-(body - lead) // rest)
        per = -(-per // 256) * 256
        return (1 if lead else 0) + -(-(body - lead) // per) + (1 if tail else 0)

    @property
    def step_index(self):
        return int(self.lib.g2048_env_step_index(self._h))

    def close(self):
        if self._h:
            self.lib.g2048_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StepSchedule:
    """A pre-built list of step() calls, issued by ONE C call (g2048_step_list) instead of one trip through the
    interpreter per launch.

        sched = StepSchedule()
        for j in range(K):
            sched.add(games[j % len(games)], actions[j])       # any mix of BatchedGame2048 objects on ONE device
        sched.run()                                            # K kernel launches, back to back

    Each add() records a complete g2048_step call — the game's buffers, the action row and the step index that
    call will have — and advances the game's host-side step index, exactly as game.step() would; run() launches
    the recorded calls in order on the current stream.  A schedule is a one-shot object: the step indices are
    baked in, so run() (or run(lo, hi) over consecutive slices) must be executed exactly once and in order.
    Results land in each game's own output tensors (rewards, dones, ...), overwritten by that game's next step.
    Use: many small vectorised envs stepped round-robin, or open-loop rollouts of small batches, where the GPU
    finishes a step (~3 us for 131,072 boards) faster than Python can issue the next one (~6 us)."""

    def __init__(self):
        self._items = []
        self._owner = []         # id() of the game of every item (run_threads deals games to issuing threads)
        self._keep = []          # action tensors must outlive the launches
        self._array = None
        self._device = None
        self._next = 0

    def __len__(self):
        return len(self._items)

    def add(self, game, actions=None, policy=None, chained=False):
        """Record game.step(actions) — or game.step(policy=...), the actions then being drawn by the kernel and written
        to `actions` (or the game's internal buffer).  chained: as in BatchedGame2048.step (the action rows of a
        schedule exist before it runs, so a schedule that is the only thing stepping its games may chain them all)."""
        n = game.num_envs
        pol = 0
        if policy is not None:
            if policy not in ("uniform", "legal"):
                raise ValueError("policy must be None, 'uniform' or 'legal'")
            if policy == "legal" and game.legal_mask is None:
                raise ValueError("policy='legal' needs the 'legal_mask' output")
            pol = _lib.FLAG_POLICY_LEGAL if policy == "legal" else _lib.FLAG_POLICY_UNIFORM
            if actions is None:
                if game._actions_out is None:
                    game._actions_out = torch.empty(n, dtype=torch.uint8, device=game.device)
                actions = game._actions_out
        if not (isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8 and actions.device == game.device
                and actions.is_contiguous() and actions.shape == (n,)):
            raise ValueError("actions must be a contiguous uint8 [%d] tensor on %s" % (n, game.device))
        if game._step_counter is not None:
            raise G2048Error("StepSchedule needs the host-side step index (no device step counter)")
        if self._device is None:
            self._device = game.device
        elif self._device != game.device:
            raise ValueError("all games of a schedule must live on one device")
        if self._array is not None:
            raise G2048Error("the schedule was already built; make a new one")
        game._build_step_args()
        a = StepArgs.from_buffer_copy(game._args)
        a.actions = actions.data_ptr()
        a.step_index = game.step_index
        a.boards = game.boards.data_ptr()
        a.flags = (FLAG_AUTO_RESET if game.auto_reset else 0) | pol
        game._chain_args(a, chained)
        game.step_index += 1
        self._items.append(a)
        self._keep.append((game, actions))
        self._owner.append(id(game))

    def build(self):
        if self._array is None:
            arr = (StepArgs * len(self._items))(*self._items)
            self._array, self._lib = arr, _lib.lib()
        return self._array

    def run(self, lo=None, hi=None, events=None, every=0):
        """Launch items [lo, hi) (default: everything not yet launched) on the current stream.
        events (a list of torch.cuda.Event(enable_timing=True)) with every = K: events[0] is recorded before the
        first launch and events[r] after launch r*K, all from C (g2048_step_list_timed) — R timed regions of K
        launches cost one trip through the interpreter instead of R."""
        arr = self.build()
        lo = self._next if lo is None else int(lo)
        hi = len(self._items) if hi is None else int(hi)
        if lo != self._next or hi < lo or hi > len(self._items):
            raise G2048Error("a StepSchedule runs once, in order: next item is %d, got [%d, %d)" % (self._next, lo, hi))
        if hi == lo:
            return
        idx = self._device.index
        ptr = C.c_void_p(C.addressof(arr) + lo * C.sizeof(StepArgs))

        def call():
            if not events:
                return self._lib.g2048_step_list(ptr, hi - lo, _raw_stream(idx))
            for e in events:                       # torch creates the cudaEvent_t lazily, at the first record()
                if not e.cuda_event:
                    e.record()
            handles = (C.c_void_p * len(events))(*[e.cuda_event for e in events])
            return self._lib.g2048_step_list_timed(ptr, hi - lo, int(every), handles, len(events), _raw_stream(idx))
        if torch.cuda.current_device() == idx:
            rc = call()
        else:
            with torch.cuda.device(idx):
                rc = call()
        self._next = hi
        if rc:
            check(rc)

    def run_threads(self, n_threads=2):
        """Launch everything not yet launched from `n_threads` host threads on as many streams.  One host thread gets
        a kernel launch out every ~2-2.4 us, two ~1.4 us between them (scripts/micro/launch_cost.cu); a B200 steps
        131,072 boards in ~1.3 us when chained launches of several env sets share it.  The games are dealt to the
        threads round-robin in order of first appearance; the launches of one game stay in order on one stream (so
        chained launches remain valid), launches of different games are unordered among themselves — as they may be:
        they share nothing.  The side streams start after the work queued on the current stream and the current
        stream continues after them."""
        import threading
        n_threads = int(n_threads)
        arr = self.build()
        lo, hi = self._next, len(self._items)
        if hi == lo:
            return
        if n_threads <= 1:
            return self.run()
        order, parts = {}, [[] for _ in range(n_threads)]
        for j in range(lo, hi):
            t = order.setdefault(self._owner[j], len(order) % n_threads)
            parts[t].append(j)
        parts = [p_ for p_ in parts if p_]
        arrays = [(StepArgs * len(p_))(*[arr[j] for j in p_]) for p_ in parts]
        dev = self._device
        cur = torch.cuda.current_stream(dev)
        streams = [torch.cuda.Stream(device=dev) for _ in parts]
        rcs = [0] * len(parts)

        def issue(t):
            with torch.cuda.device(dev):
                streams[t].wait_stream(cur)
                rcs[t] = self._lib.g2048_step_list(arrays[t], len(parts[t]), C.c_void_p(streams[t].cuda_stream))
        threads = [threading.Thread(target=issue, args=(t,)) for t in range(len(parts))]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        for st in streams:
            cur.wait_stream(st)
        self._next = hi
        for rc in rcs:
            if rc:
                check(rc)

