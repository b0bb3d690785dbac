"""GPU tests of the callers of the step path: the on-device rollout
collector (buffer chain replayed on the oracle, GAE kernel against a plain-torch float32
reference of the same recurrence) and the batched evaluator (per-episode results against the
same evaluation run on the oracle with the same deterministic policy)."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gym_2048_b200 as g
    return g


def test_gae_kernel_equals_torch_float32_reference(G):
    import ctypes as C
    import torch
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    for T, n in ((1, 5), (7, 1), (64, 1000), (256, 4099)):
        rewards = torch.randint(0, 64, (T, n), generator=gen, device=dev).float() * 4
        values = torch.randn((T, n), generator=gen, device=dev) * 30
        starts = (torch.rand((T, n), generator=gen, device=dev) < 0.07).to(torch.uint8)
        last_v = torch.randn(n, generator=gen, device=dev) * 30
        last_d = (torch.rand(n, generator=gen, device=dev) < 0.07).to(torch.uint8)
        adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
        p = lambda t: C.c_void_p(t.data_ptr())                               # noqa: E731
        rc = G._lib.lib().g2048_gae(p(rewards), p(values), p(starts), p(last_v), p(last_d), p(adv), p(ret), T, n,
                                    0.99, 0.95, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        ref_adv, ref_ret = G.gae_reference(rewards, values, starts, last_v, last_d, 0.99, 0.95)
        # float32, same operation order, no fma: tolerance 0 (bit-exact); kept as an explicit bar
        torch.testing.assert_close(adv, ref_adv, rtol=0, atol=0)
        torch.testing.assert_close(ret, ref_ret, rtol=0, atol=0)


def test_rollout_buffer_chain_replays_on_the_oracle(G):
    import torch
    torch.manual_seed(42)
    n, T = 1500, 40
    game = G.BatchedGame2048(n, seed=42, env_id_base=77)
    game.reset()
    warm = torch.randint(0, 4, (n,), device=game.device, dtype=torch.uint8)
    game.step(warm)                                              # collect() need not start at step 0
    policy = G.ResNetActorCritic(filters=8, residual_blocks=1).to(game.device).eval()
    col = G.RolloutCollector(game, policy, T, seed=5)
    ref = oracle.OracleBatch(n, seed=42, env_id_base=77)
    ref.boards[:] = game.boards.cpu().numpy()
    ref.step_index = game.step_index
    col.collect()
    b = col.boards.cpu().numpy()
    acts, rew, starts = col.actions.cpu().numpy(), col.rewards.cpu().numpy(), col.episode_starts.cpu().numpy()
    assert np.array_equal(b[0], ref.boards)
    prev_done = np.ones(n, np.uint8)
    for t in range(T):
        assert np.array_equal(starts[t], prev_done), t
        o = ref.step(acts[t])
        assert np.array_equal(b[t + 1], ref.boards), t
        assert np.array_equal(rew[t], o["rewards"]), t
        prev_done = o["dones"]
    assert np.array_equal(col.last_dones.cpu().numpy(), prev_done)
    assert game.boards.data_ptr() == col.boards[T].data_ptr()
    # advantages = the float32 reference of the recurrence on the collected buffers
    adv, ret = G.gae_reference(col.rewards, col.values, col.episode_starts, col.last_values, col.last_dones,
                               col.gamma, col.gae_lambda)
    torch.testing.assert_close(col.advantages, adv, rtol=0, atol=0)
    torch.testing.assert_close(col.returns, ret, rtol=0, atol=0)
    # sampled actions follow the policy: log_probs are the log-softmax entries of the stored actions
    obs = game.observe(torch.float32, boards=col.boards[3].contiguous())
    logits, value = policy(obs)
    lp = torch.log_softmax(logits, -1).gather(1, col.actions[3].long()[:, None]).squeeze(1)
    torch.testing.assert_close(lp, col.log_probs[3], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(value, col.values[3], rtol=1e-5, atol=1e-5)
    # minibatches re-encode the stored boards
    seen = 0
    for obs_mb, a_mb, v_mb, lp_mb, adv_mb, ret_mb in col.minibatches(4096, shuffle=False):
        assert obs_mb.shape[1:] == (16, 4, 4) and obs_mb.shape[0] == a_mb.shape[0] == adv_mb.shape[0]
        if seen == 0:
            exp = oracle.encode_obs_u8(b[:T].reshape(-1, 16)[:obs_mb.shape[0]])
            assert np.array_equal(obs_mb.cpu().numpy().astype(np.uint8), exp)
        seen += obs_mb.shape[0]
    assert seen == T * n
    # a second rollout continues from the last boards
    col.collect()
    assert np.array_equal(col.boards[0].cpu().numpy(), ref.boards)


class _IntPolicy:
    """Deterministic scores from small-integer weights: exact in float32 on any device."""

    def __init__(self, seed=0):
        rng = np.random.default_rng(seed)
        self.w = rng.integers(-3, 4, (4, 16 * 16)).astype(np.float32)
        self.bias = np.arange(4, dtype=np.float32) * 0.125          # no ties

    def numpy(self, obs_u8):
        return obs_u8.reshape(len(obs_u8), -1).astype(np.float32) @ self.w.T + self.bias

    def torch(self, dev):
        import torch
        w, b = torch.from_numpy(self.w).to(dev), torch.from_numpy(self.bias).to(dev)
        return lambda obs: obs.reshape(obs.shape[0], -1).float() @ w.T + b


@pytest.mark.parametrize("mask_illegal", [False, True])
def test_evaluate_model_matches_the_same_evaluation_on_the_oracle(G, mask_illegal):
    import torch
    episodes, seed = 300, 456
    pol = _IntPolicy(1)
    res = G.evaluate_model(pol.torch(torch.device("cuda", 0)), episodes, epsilon=0.0, seed=seed,
                           mask_illegal=mask_illegal, max_moves=150 if mask_illegal else 2000)
    # the same loop on the CPU oracle (train.py:122-214 semantics)
    ref = oracle.OracleBatch(episodes, seed=seed, illegal_move_reward=-1.0, auto_reset=False)
    ref.reset()
    active = np.ones(episodes, bool)
    tot, moves, ill, high = np.zeros(episodes), np.zeros(episodes, int), np.zeros(episodes, int), np.zeros(episodes, int)
    cap = 150 if mask_illegal else 2000
    for _ in range(cap + 1):
        s = pol.numpy(oracle.encode_obs_u8(ref.boards))
        if mask_illegal:
            lm = oracle.status(ref.boards)["legal_mask"]
            legal = ((lm[:, None] >> np.arange(4)[None, :]) & 1).astype(bool)
            legal |= ~legal.any(axis=1, keepdims=True)
            s = np.where(legal, s, -np.inf)
        o = ref.step(np.argmax(s, axis=1).astype(np.uint8))
        tot += np.where(active, o["rewards"], 0)
        ill += active & (o["illegal"] != 0)
        moves += active
        high = np.where(active, o["highest_exp"], high)
        active &= o["dones"] == 0
        if not active.any():
            break
    eps = res["Episodes"]
    assert [e["total_reward"] for e in eps] == tot.tolist()
    assert [e["moves"] for e in eps] == moves.tolist()
    assert [e["illegal_moves"] for e in eps] == ill.tolist()
    assert [e["highest"] for e in eps] == [int(1 << h) if h else 0 for h in high]
    assert res["Average score"] == sum(tot.tolist()) / episodes and res["Max score"] == tot.max()
    assert res["Highest tile"] == int(1 << high.max())
    if not mask_illegal:
        assert max(ill) == 1 and min(ill) >= 0                    # an illegal move ends the episode (:91-95)
    else:
        assert moves.max() == cap + 1 or not active.any()


def test_evaluate_report_csv_and_epsilon(G, tmp_path, monkeypatch):
    import torch
    monkeypatch.chdir(tmp_path)
    pol = _IntPolicy(2)
    res = G.evaluate_model(pol.torch(torch.device("cuda", 0)), 64, epsilon=1.0, seed=456, agent_seed=123)
    assert len(res["Episodes"]) == 64 and all(e["moves"] >= 1 for e in res["Episodes"])
    G.report_evaluation_results(res, label="t")
    lines = open(tmp_path / "scores_t.csv").read().split("\n")
    assert lines[0] == "total_reward,highest,moves,illegal_moves" and len(lines) == 66
    e0 = res["Episodes"][0]
    assert lines[1] == "%s,%d,%d,%d" % (e0["total_reward"], e0["highest"], e0["moves"], e0["illegal_moves"])
    # epsilon = 1: every action is random; same agent seed -> same result, another seed -> different
    again = G.evaluate_model(pol.torch(torch.device("cuda", 0)), 64, epsilon=1.0, seed=456, agent_seed=123)
    other = G.evaluate_model(pol.torch(torch.device("cuda", 0)), 64, epsilon=1.0, seed=456, agent_seed=124)
    assert again["Episodes"] == res["Episodes"] and other["Episodes"] != res["Episodes"]
