"""Chained launches (G2048StepArgs.chain, G2048_FLAG_CHAINED / _CHAIN_INTERLEAVED, include/g2048.h): a step that
depends on its predecessor warp by warp instead of grid by grid must leave exactly what plain launches leave —
boards, rewards, dones, legal masks, and the running episode statistics, which fold EVERY step of a rollout into
the final state (read-modify-write arrays: a launch that ran ahead of its predecessor would corrupt them).  The
launches are issued back to back (no reads in between), so that consecutive launches really overlap on the GPU.
The plain path these are compared with is itself checked against the oracle and the reference's golden vectors
(tests/test_gpu_parity.py); test_step_n_is_k_calls_of_step there runs g2048_step_n, which chains by default."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

ALL = ("illegal", "highest", "legal_mask", "terminal", "episode")


def _acts(K, n, seed=1):
    import torch
    return torch.randint(0, 4, (K, n), device="cuda", dtype=torch.uint8, generator=torch.Generator(device="cuda").manual_seed(seed))


def _same_state(a, b):
    import torch
    assert torch.equal(a.boards, b.boards)
    assert torch.equal(a.rewards, b.rewards) and torch.equal(a._dones, b._dones)
    assert a.step_index == b.step_index
    for name in ("legal_mask", "highest_exp", "_illegal", "ep_score", "ep_len", "ep_return", "final_score", "final_len", "final_return"):
        x, y = getattr(a, name), getattr(b, name)
        assert (x is None) == (y is None)
        if x is not None:
            assert torch.equal(x, y), name


@pytest.mark.parametrize("n,K,outputs,mode", [
    (1, 5, (), True),
    (1000, 40, ALL, True),
    (70001, 60, ("episode",), True),              # one board per thread, ragged last CTA
    (300000, 60, ("episode",), True),             # two loop iterations per thread in the direct shape (592 x 256 threads)
    (300000, 60, ("episode",), "interleaved"),    # eight in the interleaved one (148 x 256)
    (1 << 18, 50, ("legal_mask", "episode"), True),
    (1 << 20, 24, (), "interleaved"),
])
def test_chained_steps_leave_what_plain_steps_leave(n, K, outputs, mode):
    import torch
    import gym_2048_b200 as g
    mk = lambda: g.BatchedGame2048(n, seed=9, env_id_base=12345, outputs=outputs, illegal_move_reward=-1.0)   # noqa: E731
    a, b = mk(), mk()
    a.reset(), b.reset()
    acts = _acts(K, n)
    torch.cuda.synchronize()
    for k in range(K):
        a.step(acts[k], chained=mode)
    for k in range(K):
        b.step(acts[k])
    _same_state(a, b)
    ref = oracle.OracleBatch(n, seed=9, env_id_base=12345, illegal_move_reward=-1.0, threads=8)
    ref.reset()
    for k in range(K):
        ref.step(acts[k].cpu().numpy())
    assert np.array_equal(a.boards.cpu().numpy(), ref.boards)


@pytest.mark.parametrize("policy,outputs", [("uniform", ()), ("legal", ("legal_mask", "episode"))])
def test_chained_in_kernel_policy(policy, outputs):
    """BASELINE config 4 as chained launches: the kernel draws among the legal moves of the mask its predecessor wrote."""
    import torch
    import gym_2048_b200 as g
    n, K = 1 << 18, 80
    mk = lambda: g.BatchedGame2048(n, seed=5, outputs=outputs)      # noqa: E731
    a, b = mk(), mk()
    a.reset(), b.reset()
    for k in range(K):
        a.step(policy=policy, chained=True)
    for k in range(K):
        b.step(policy=policy)
    _same_state(a, b)
    assert torch.equal(a._actions_out, b._actions_out)


def test_interleaved_chains_round_robin_schedule():
    """Several env sets stepped round-robin by one C call, every launch chained to its own set's previous one."""
    import torch
    import gym_2048_b200 as g
    n, K, S = 200000, 90, 3
    mk = lambda s: g.BatchedGame2048(n, seed=3, env_id_base=s * n, outputs=("episode",) if s == 1 else ())   # noqa: E731
    A, B = [mk(s) for s in range(S)], [mk(s) for s in range(S)]
    for x in A + B:
        x.reset()
    acts = _acts(16, n, seed=2)
    sched = g.StepSchedule()
    for j in range(K):
        sched.add(A[j % S], acts[j % 16], chained="interleaved")
    sched.run()
    for j in range(K):
        B[j % S].step(acts[j % 16])
    for a, b in zip(A, B):
        _same_state(a, b)


def test_schedule_issued_by_two_host_threads_on_two_streams():
    """StepSchedule.run_threads: the games dealt to two issuing threads / streams, chained launches — the same boards
    as the same steps made one by one."""
    import torch
    import gym_2048_b200 as g
    n, K, S = 131072, 400, 4
    mk = lambda s: g.BatchedGame2048(n, seed=13, env_id_base=s * n, outputs=("episode",) if s % 2 else ())   # noqa: E731
    A, B = [mk(s) for s in range(S)], [mk(s) for s in range(S)]
    for x in A + B:
        x.reset()
    acts = _acts(16, n, seed=3)
    sched = g.StepSchedule()
    for j in range(K):
        sched.add(A[j % S], acts[j % 16], chained="interleaved")
    sched.run(0, 40)                      # a first slice from the calling thread, the rest from two threads
    sched.run_threads(2)
    for j in range(K):
        B[j % S].step(acts[j % 16])
    torch.cuda.synchronize()
    for a, b in zip(A, B):
        _same_state(a, b)


def test_chained_and_plain_env_sets_share_a_stream():
    """A chained launch does not wait for the previous grid before its work; stream order must still hold for what
    follows it: a PLAIN env stepped in between (programmatic dependent launches that wait for "the previous grid")
    sees its own previous step complete."""
    import torch
    import gym_2048_b200 as g
    n, K = 1 << 18, 60
    mk = lambda s, outs: g.BatchedGame2048(n, seed=8, env_id_base=s * n, outputs=outs)      # noqa: E731
    a1, c1, a2, c2 = mk(0, ("episode",)), mk(1, ("episode",)), mk(0, ("episode",)), mk(1, ("episode",))
    for x in (a1, c1, a2, c2):
        x.reset()
    acts = _acts(8, n, seed=4)
    sched = g.StepSchedule()
    for k in range(K):
        sched.add(a1, acts[k % 8], chained="interleaved")
        sched.add(c1, acts[(k + 3) % 8])                      # plain: no chain buffer at all
    sched.run()
    for k in range(K):
        a2.step(acts[k % 8])
    for k in range(K):
        c2.step(acts[(k + 3) % 8])
    _same_state(a1, a2)
    _same_state(c1, c2)


def test_a_chain_survives_everything_else_the_class_does():
    """reset(mask), set_boards, step_many, a plain step and a change of chaining mode between chained steps: the
    class issues the next step unchained; results equal an env that never chains."""
    import torch
    import gym_2048_b200 as g
    n = 50000
    mk = lambda: g.BatchedGame2048(n, seed=21, outputs=("legal_mask", "illegal", "highest"))       # noqa: E731
    a, b = mk(), mk()
    a.reset(), b.reset()
    acts = _acts(64, n, seed=6)
    mask = (torch.arange(n, device="cuda") % 3 == 0).to(torch.uint8)

    def both(fn):
        fn(a, True), fn(b, False)
    t = 0
    for rnd in range(4):
        for _ in range(5):
            both(lambda e, ch: e.step(acts[t % 64], chained=ch and (True if rnd % 2 == 0 else "interleaved")))
            t += 1
        if rnd == 0:
            both(lambda e, ch: e.reset(mask=mask))
        elif rnd == 1:
            saved = b.boards.clone()
            both(lambda e, ch: e.set_boards(saved.flip(0)))
        elif rnd == 2:
            both(lambda e, ch: e.step_many(acts[:7].contiguous()))
        both(lambda e, ch: e.step(acts[(t + 1) % 64]))           # a plain step in between
    _same_state(a, b)


def test_chain_argument_checks_and_raw_abi():
    """The C ABI directly: a zeroed chain buffer, step_index advancing by one; flag without buffer and a misaligned
    buffer are refused; a launch at step_index 0 is never chained."""
    import torch
    import gym_2048_b200 as g
    L = g._lib.lib()
    n, K = 40000, 30
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    boards = [torch.zeros((n, 16), dtype=torch.uint8, device=dev) for _ in range(2)]
    rew = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2)]
    done = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    chain = torch.zeros(g._lib.CHAIN_WORDS, dtype=torch.int64, device=dev)
    acts = _acts(K, n, seed=11)
    for s in range(2):
        assert L.g2048_reset(boards[s].data_ptr(), None, n, 0, 77, 0, stream) == 0
    for s in range(2):
        a = g._lib.StepArgs()
        a.boards, a.rewards, a.dones, a.n, a.seed = boards[s].data_ptr(), rew[s].data_ptr(), done[s].data_ptr(), n, 77
        for k in range(K):
            a.actions, a.step_index = acts[k].data_ptr(), k
            a.flags = g._lib.FLAG_AUTO_RESET | (g._lib.FLAG_CHAINED if s == 0 else 0)      # k = 0: treated as unchained
            a.chain = chain.data_ptr() if s == 0 else None
            assert L.g2048_step(C.byref(a), stream) == 0, L.g2048_last_error()
    assert torch.equal(boards[0], boards[1]) and torch.equal(rew[0], rew[1]) and torch.equal(done[0], done[1])
    words = chain.cpu().numpy()
    assert set(np.unique(words)) == {0, K}            # every warp of the launch shape published the last step index + 1
    a.chain, a.flags = None, g._lib.FLAG_CHAINED
    assert L.g2048_step(C.byref(a), stream) == -1 and b"chain buffer" in L.g2048_last_error()
    a.chain, a.flags = chain.data_ptr() + 4, 0
    assert L.g2048_step(C.byref(a), stream) == -2 and b"8-byte aligned" in L.g2048_last_error()
