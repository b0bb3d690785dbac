"""On-device PPO rollout collector — the caller side of the step path (SURVEY §8f row 2,
BASELINE config 5).

Replaces SB3's `OnPolicyAlgorithm.collect_rollouts` + `RolloutBuffer` as driven by
`/root/reference/ppo_train.py:138-183` (n_envs Python env objects stepped one by one, float32
observations of 1 KiB per env-step kept in host memory) with a loop that never leaves the GPU:
one-hot encode -> policy -> sample -> g2048_step, the boards stepped OUT OF PLACE straight into
the rollout buffer.  The buffer keeps the 16-byte boards, not the observations (64x smaller
than SB3's float32 obs) and re-encodes them per minibatch; advantages come from the
g2048_gae kernel.  No host round trip inside collect()."""
import ctypes as C

import torch

from ._lib import check


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class RolloutCollector:
    def __init__(self, game, policy, horizon, gamma=0.99, gae_lambda=0.95, obs_dtype=torch.float32,
                 channels_last=False, seed=0):
        self.game, self.policy, self.T = game, policy, int(horizon)
        self.gamma, self.gae_lambda = float(gamma), float(gae_lambda)
        n, dev = game.num_envs, game.device
        self.n, self.device = n, dev
        T = self.T
        self.boards = torch.zeros((T + 1, n, 16), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((T, n), dtype=torch.uint8, device=dev)
        self.rewards = torch.zeros((T, n), dtype=torch.float32, device=dev)
        self.episode_starts = torch.zeros((T, n), dtype=torch.uint8, device=dev)
        self.values = torch.zeros((T, n), dtype=torch.float32, device=dev)
        self.log_probs = torch.zeros((T, n), dtype=torch.float32, device=dev)
        self.advantages = torch.zeros((T, n), dtype=torch.float32, device=dev)
        self.returns = torch.zeros((T, n), dtype=torch.float32, device=dev)
        self.last_values = torch.zeros(n, dtype=torch.float32, device=dev)
        self.last_dones = torch.ones(n, dtype=torch.uint8, device=dev)     # a fresh env starts an episode
        self.obs_dtype = obs_dtype
        self.channels_last = channels_last
        self._obs = torch.empty((n, 16, 4, 4), dtype=obs_dtype, device=dev)
        self.gen = torch.Generator(device=dev).manual_seed(int(seed))
        self.env_steps = 0

    def _policy_obs(self, obs):
        return obs.contiguous(memory_format=torch.channels_last) if self.channels_last else obs

    @torch.no_grad()
    def collect(self):
        """One rollout of `horizon` steps for every env; fills the buffers and the advantages."""
        g, T = self.game, self.T
        self.boards[0].copy_(g.boards)
        g.boards = self.boards[0]
        for t in range(T):
            obs = g.observe(self.obs_dtype, out=self._obs)
            logits, value = self.policy(self._policy_obs(obs))
            logp_all = torch.log_softmax(logits.float(), dim=-1)
            # Gumbel-max sampling: argmax(logp + G) ~ Categorical(softmax(logits))
            u = torch.rand(logp_all.shape, generator=self.gen, device=self.device).clamp_(1e-20, 1.0)
            act = torch.argmax(logp_all - torch.log(-torch.log(u)), dim=-1)
            self.actions[t].copy_(act)
            self.log_probs[t].copy_(logp_all.gather(1, act[:, None]).squeeze(1))
            self.values[t].copy_(value.float())
            self.episode_starts[t].copy_(self.last_dones)
            res = g.step(self.actions[t], boards_out=self.boards[t + 1])
            self.rewards[t].copy_(res.rewards)
            self.last_dones.copy_(res.dones)
        obs = g.observe(self.obs_dtype, out=self._obs)
        _, value = self.policy(self._policy_obs(obs))
        self.last_values.copy_(value.float())
        self.compute_returns_and_advantage()
        self.env_steps += T * self.n
        return self

    def compute_returns_and_advantage(self):
        with torch.cuda.device(self.device):
            check(self.game.lib.g2048_gae(_ptr(self.rewards), _ptr(self.values), _ptr(self.episode_starts),
                                          _ptr(self.last_values), _ptr(self.last_dones), _ptr(self.advantages),
                                          _ptr(self.returns), self.T, self.n, self.gamma, self.gae_lambda,
                                          C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def minibatches(self, batch_size, shuffle=True):
        """SB3 RolloutBuffer.get(): (obs, actions, old_values, old_log_prob, advantages, returns),
        observations re-encoded from the stored boards."""
        total = self.T * self.n
        idx = torch.randperm(total, generator=self.gen, device=self.device) if shuffle else \
            torch.arange(total, device=self.device)
        flat_boards = self.boards[:self.T].reshape(total, 16)
        flat = lambda x: x.reshape(total)                                         # noqa: E731
        for lo in range(0, total, batch_size):
            sel = idx[lo:lo + batch_size]
            obs = self.game.observe(self.obs_dtype, boards=flat_boards[sel].contiguous())
            yield (obs, flat(self.actions)[sel].long(), flat(self.values)[sel], flat(self.log_probs)[sel],
                   flat(self.advantages)[sel], flat(self.returns)[sel])


def gae_reference(rewards, values, episode_starts, last_values, last_dones, gamma, gae_lambda):
    """Plain-torch float32 restatement of SB3's compute_returns_and_advantage (the numerics
    reference for the g2048_gae kernel; used by tests)."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(last_values)
    g32 = torch.tensor(gamma, dtype=torch.float32, device=rewards.device)
    gl32 = torch.tensor(gamma * gae_lambda, dtype=torch.float32, device=rewards.device)
    for t in reversed(range(T)):
        if t == T - 1:
            nnt, nv = 1.0 - last_dones.float(), last_values
        else:
            nnt, nv = 1.0 - episode_starts[t + 1].float(), values[t + 1]
        delta = rewards[t] + g32 * nv * nnt - values[t]
        last = delta + gl32 * nnt * last
        adv[t] = last
    return adv, adv + values
