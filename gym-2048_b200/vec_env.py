"""Game2048VecEnv — Stable-Baselines3 `VecEnv`-shaped adapter over BatchedGame2048.

Replaces `make_vec_env("2048-v0", n_envs)` at `/root/reference/ppo_train.py:123` (SB3's
DummyVecEnv + Monitor: a Python loop over env objects) with one kernel launch per step.
Semantics kept from SB3: `step_wait()` -> (obs, rewards float32, dones bool, infos), same-step
auto-reset with `infos[i]["terminal_observation"]`, `["TimeLimit.truncated"] = False`,
`["episode"] = {"r", "l", "t"}` at episode end (`r` = the sum of the episode's REWARDS, illegal-move penalty
included, as Monitor computes it; the game score — merge scores only — is `["score"]`), plus the reference's
`["highest"]` and `["illegal_move"]` (ppo_train.py:76-82 reads `infos[i]["highest"]` on done).  Info dicts are
materialised only for envs that finished; all others share one read-only empty-ish dict, so
the per-step Python cost does not grow with the batch.

SB3 itself is not a dependency: the class follows its method names so `PPO("CnnPolicy",
Game2048VecEnv(...))` works where SB3 is installed (it subclasses VecEnv there), and the
restated rollout loop in bench/tests works where it is not.
"""
import time

import numpy as np
import torch

from .batched import BatchedGame2048

try:  # pragma: no cover - SB3 is not in this image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except ImportError:
    _VecEnvBase = object

try:
    from gymnasium import spaces as _spaces
except ImportError:
    _spaces = None

from .env import _box, _discrete


class Game2048VecEnv(_VecEnvBase):
    def __init__(self, num_envs, seed=0, device=None, obs_dtype=torch.float32, return_torch=False,
                 illegal_move_reward=0.0, max_tile=None, env_id_base=0):
        self.game = BatchedGame2048(num_envs, seed=seed, device=device, env_id_base=env_id_base,
                                    illegal_move_reward=illegal_move_reward, max_tile=max_tile, auto_reset=True,
                                    outputs=("illegal", "highest", "legal_mask", "episode", "terminal"))
        self.num_envs = int(num_envs)
        self.observation_space = _box(0, 1, (16, 4, 4), int)          # reference :51-52
        self.action_space = _discrete(4)                              # reference :49
        self.obs_dtype = obs_dtype
        self.return_torch = return_torch
        self.render_mode = None
        self._actions = None
        self._t0 = time.time()
        self._shared_info = {"TimeLimit.truncated": False}
        if _VecEnvBase is not object:
            super().__init__(self.num_envs, self.observation_space, self.action_space)

    # -- VecEnv protocol -----------------------------------------------------------------
    def _obs(self):
        obs = self.game.observe(self.obs_dtype)
        return obs if self.return_torch else obs.cpu().numpy()

    def seed(self, seed=None):
        if seed is not None:
            self.game.reset(seed=seed)
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def reset(self):
        self.game.reset()
        return self._obs()

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        g = self.game
        r = g.step(self._actions)
        obs = self._obs()
        dones = r.dones
        infos = [self._shared_info] * self.num_envs
        idx = torch.nonzero(dones).flatten()
        if idx.numel():
            idx_c = idx.cpu().numpy()
            term_obs = g.observe(self.obs_dtype, boards=r.terminal_boards[idx].contiguous())
            term_obs = term_obs if self.return_torch else term_obs.cpu().numpy()
            fr = r.final_return[idx].cpu().numpy()       # Monitor's 'r': the sum of the rewards the agent saw
            fs = r.final_score[idx].cpu().numpy()        # the game score (merge scores only)
            fl = r.final_len[idx].cpu().numpy()
            hi = r.highest_exp[idx].cpu().numpy().astype(np.int64)
            il = r.illegal[idx].cpu().numpy()
            now = round(time.time() - self._t0, 6)
            for j, i in enumerate(idx_c):
                infos[i] = {
                    "TimeLimit.truncated": False,
                    "terminal_observation": term_obs[j],
                    "episode": {"r": round(float(fr[j]), 6), "l": int(fl[j]), "t": now},
                    "score": int(fs[j]),
                    "highest": int(1 << hi[j]) if hi[j] else 0,
                    "illegal_move": bool(il[j]),
                }
        if self.return_torch:
            return obs, r.rewards.clone(), dones.clone(), infos
        return obs, r.rewards.cpu().numpy(), dones.cpu().numpy(), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        pass

    # -- SB3 attribute plumbing ------------------------------------------------------------
    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(list(indices)) if not isinstance(indices, int) else 1
        return [getattr(self.game, attr_name, getattr(self, attr_name, None))] * n

    def set_attr(self, attr_name, value, indices=None):
        if attr_name == "illegal_move_reward":
            self.game.set_illegal_move_reward(value)
        elif attr_name == "max_tile":
            self.game.set_max_tile(value)
        else:
            setattr(self, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        if method_name in ("set_illegal_move_reward", "set_max_tile"):
            return [getattr(self.game, method_name)(*method_args, **method_kwargs)] * self.num_envs
        raise NotImplementedError(method_name)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def get_images(self):
        return [None] * self.num_envs

    # -- extras ----------------------------------------------------------------------------
    def action_masks(self):
        """bool [n,4]: legal moves of the boards just returned (sb3-contrib MaskablePPO hook)."""
        m = self.game.legal_mask
        out = ((m[:, None] >> torch.arange(4, device=m.device, dtype=torch.uint8)[None, :]) & 1).bool()
        return out if self.return_torch else out.cpu().numpy()
