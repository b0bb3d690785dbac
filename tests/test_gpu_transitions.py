"""GPU parity of the rows either side of the step: out-of-place stepping
into a trajectory buffer, the device random policies, the transition recorder, board
symmetries / augmentation / discounted return (fixtures from the UNMODIFIED reference's
training_data.py, plus the oracle on large random inputs) and the CSV schema, all through the
C ABI."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gym_2048_b200 as g
    return g


@pytest.fixture(scope="module")
def gold():
    return load_golden("transitions.npz")


def _t(G, gold, name):
    return G.Transitions(gold[name + "/in_x"], gold[name + "/in_y"], gold[name + "/in_reward"].astype(np.float32),
                         gold[name + "/in_next_x"], gold[name + "/in_done"])


def _same(t, gold, prefix, full=False):
    assert np.array_equal(t.boards.cpu().numpy(), gold[prefix + "_x"])
    assert np.array_equal(t.next_boards.cpu().numpy(), gold[prefix + "_next_x"])
    assert np.array_equal(t.actions.cpu().numpy(), gold[prefix + "_y"])
    if full:
        assert np.array_equal(t.rewards.cpu().numpy().astype(np.float64), gold[prefix + "_reward"])
        assert np.array_equal(t.dones.cpu().numpy(), gold[prefix + "_done"])


@pytest.mark.parametrize("name", ["csv", "syn"])
def test_symmetries_and_augment_match_reference(G, gold, name):
    t = _t(G, gold, name); t.hflip(); _same(t, gold, name + "/hflip")
    for k in (1, 2, 3):
        t = _t(G, gold, name); t.rotate(k); _same(t, gold, "%s/rot%d" % (name, k))
    t = _t(G, gold, name); t.hflip(); t.rotate(3); _same(t, gold, name + "/hflip_rot3")
    t = _t(G, gold, name); t.augment(); _same(t, gold, name + "/aug", full=True)
    assert t.size() == 8 * len(gold[name + "/in_y"])


@pytest.mark.parametrize("name", ["csv", "syn"])
def test_discounted_return_bit_exact_vs_reference(G, gold, name):
    t = _t(G, gold, name)
    assert np.array_equal(t.get_discounted_return().cpu().numpy().reshape(-1), gold[name + "/ret_090"])
    assert np.array_equal(t.get_discounted_return(gamma=0.99).cpu().numpy().reshape(-1), gold[name + "/ret_099"])


def test_large_random_vs_oracle(G):
    rng = np.random.default_rng(11)
    n = 300001                                                     # ragged size
    b = rng.integers(0, 19, (n, 16)).astype(np.uint8)
    nb = rng.integers(0, 19, (n, 16)).astype(np.uint8)
    a = rng.integers(0, 4, n).astype(np.uint8)
    r = rng.choice([0.0, 4.0, 8.0, 264.0, -1.0], n).astype(np.float32)
    d = (rng.random(n) < 0.05).astype(np.uint8)
    t = G.Transitions(b, a, r, nb, d)
    assert np.array_equal(t.get_discounted_return().cpu().numpy().reshape(-1), oracle.discounted_return(r, d, 0.9))
    u = t.copy(); u.augment()
    o = oracle.augment(b, nb, a, r, d)
    for k, v in (("boards", u.boards), ("next_boards", u.next_boards), ("actions", u.actions), ("rewards", u.rewards),
                 ("dones", u.dones)):
        assert np.array_equal(v.cpu().numpy(), o[k]), k
    # a long episode without any done: one thread walks the whole chain
    d0 = np.zeros(5000, np.uint8)
    t0 = G.Transitions(b[:5000], a[:5000], r[:5000], nb[:5000], d0)
    assert np.array_equal(t0.get_discounted_return(0.5).cpu().numpy().reshape(-1),
                          oracle.discounted_return(r[:5000], d0, 0.5))
    # empty table
    e = G.Transitions()
    e.augment(); e.hflip()
    assert e.size() == 0 and e.get_discounted_return().numel() == 0


def test_csv_round_trip_and_reference_bytes(G, gold, tmp_path):
    t = _t(G, gold, "syn")
    p = str(tmp_path / "t.csv")
    t.export_csv(p)
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "ref_export.csv"), "rb").read()
    t.export_csv(p, add_returns=True)
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "ref_export_returns.csv"), "rb").read()
    back = G.Transitions().import_csv(p)
    _same(back, gold, "syn/in", full=True)
    # getters present the reference's shapes and tile VALUES
    assert tuple(back.get_x().shape) == (300, 4, 4) and back.get_x().dtype.is_floating_point is False
    assert np.array_equal(back.get_x().cpu().numpy().reshape(-1, 16), oracle.exp_to_values(gold["syn/in_x"]))
    assert tuple(back.get_y_digit().shape) == (300, 1) and tuple(back.get_done().shape) == (300, 1)
    assert back.get_highest_tile() == int(oracle.exp_to_values(gold["syn/in_next_x"]).max())


@pytest.mark.parametrize("legal", [False, True])
def test_sample_actions_vs_oracle_and_sharding(G, legal):
    import torch
    n, base = 70001, (1 << 35) + 5
    game = G.BatchedGame2048(n, seed=456, env_id_base=base)
    ref = oracle.OracleBatch(n, seed=456, env_id_base=base)
    game.reset(); ref.reset()
    mask = oracle.status(ref.boards)["legal_mask"]
    for t in range(40):
        act = game.sample_actions(legal=legal)
        exp = oracle.sample_actions(mask if legal else None, n, base, 456, t)
        assert np.array_equal(act.cpu().numpy(), exp), t
        r = game.step(act)
        o = ref.step(exp)
        mask = o["legal_mask"]
        assert np.array_equal(r.boards.cpu().numpy(), ref.boards)
        if legal:
            assert not bool(r.illegal.any())                      # a legal action never is an illegal move
    torch.cuda.synchronize()


def test_out_of_place_step_equals_in_place(G):
    import torch
    n, T = 20000, 24
    a = G.BatchedGame2048(n, seed=1)
    b = G.BatchedGame2048(n, seed=1)
    a.reset(); b.reset()
    traj = torch.zeros((T + 1, n, 16), dtype=torch.uint8, device=a.device)
    term = torch.zeros((T, n, 16), dtype=torch.uint8, device=a.device)
    traj[0].copy_(b.boards)
    b.boards = traj[0]
    gen = torch.Generator(device=a.device).manual_seed(3)
    for t in range(T):
        act = torch.randint(0, 4, (n,), generator=gen, device=a.device, dtype=torch.uint8)
        before = a.boards.clone()
        ra = a.step(act)
        rb = b.step(act, boards_out=traj[t + 1], terminal_out=term[t])
        assert torch.equal(traj[t], before)                        # the input slice is left untouched
        assert torch.equal(ra.boards, rb.boards) and rb.boards.data_ptr() == traj[t + 1].data_ptr()
        assert torch.equal(ra.rewards, rb.rewards) and torch.equal(ra.dones, rb.dones)
        assert torch.equal(ra.legal_mask, rb.legal_mask) and torch.equal(ra.final_score, rb.final_score)
        d = ra.dones
        assert torch.equal(ra.terminal_boards[d], term[t][d])
    with pytest.raises(ValueError):
        b.step(act, boards_out=traj[0][:10])


def test_recorder_matches_oracle_rollout_in_game_order(G, tmp_path):
    import torch
    n, T = 257, 60
    game = G.BatchedGame2048(n, seed=42, illegal_move_reward=-1.0)
    ref = oracle.OracleBatch(n, seed=42, illegal_move_reward=-1.0)
    game.reset(); ref.reset()
    rec = G.TransitionRecorder(game, T)
    rows = [[] for _ in range(n)]
    for t in range(T):
        act = game.sample_actions()
        before = ref.boards.copy()
        o = ref.step(act.cpu().numpy())
        rec.step(act)
        for i in range(n):
            if o["illegal"][i]:
                continue                                           # gather_training_data.py:193: not recorded
            nxt = o["terminal_boards"][i] if o["dones"][i] else ref.boards[i]
            rows[i].append((before[i], int(act[i]), float(o["rewards"][i]), nxt.copy(), int(o["dones"][i])))
    tr = rec.transitions()
    flat = [r for env_rows in rows for r in env_rows]
    assert tr.size() == len(flat)
    assert np.array_equal(tr.boards.cpu().numpy(), np.stack([r[0] for r in flat]))
    assert np.array_equal(tr.actions.cpu().numpy(), np.array([r[1] for r in flat], np.uint8))
    assert np.array_equal(tr.rewards.cpu().numpy(), np.array([r[2] for r in flat], np.float32))
    assert np.array_equal(tr.next_boards.cpu().numpy(), np.stack([r[3] for r in flat]))
    assert np.array_equal(tr.dones.cpu().numpy(), np.array([r[4] for r in flat], np.uint8))
    all_rows = rec.transitions(drop_illegal=False)
    assert all_rows.size() == n * T
    with pytest.raises(IndexError):
        rec.step(act)
    # the file loads back to the same table
    p = str(tmp_path / "rec.csv")
    tr.export_csv(p, add_returns=True)
    back = G.Transitions().import_csv(p)
    assert torch.equal(back.boards, tr.boards) and torch.equal(back.next_boards, tr.next_boards)
    assert torch.equal(back.actions, tr.actions) and torch.equal(back.rewards, tr.rewards)
    # recording continues from the live boards after a rewind
    rec.rewind()
    rec.step(game.sample_actions())
    assert rec.t == 1


def test_plain_step_after_recorder_step_does_not_write_the_recorders_terminal_slice(G):
    """the cached G2048StepArgs kept `terminal_boards` pointing at the recorder's [T,n,16] slice after a
    step(terminal_out=...) on a game WITHOUT the 'terminal' output, so later plain step() calls kept writing
    terminal boards into storage the recorder may have freed.  Every per-call pointer is now reassigned per call."""
    import torch
    n = 4096
    game = G.BatchedGame2048(n, seed=5, outputs=("illegal",))          # no 'terminal' output of its own
    game.reset()
    rec = G.TransitionRecorder(game, horizon=2)
    gen = torch.Generator(device=game.device).manual_seed(3)
    rec.step(torch.randint(0, 4, (n,), generator=gen, device=game.device, dtype=torch.uint8))
    assert game._args.terminal_boards == rec.terminal[0].data_ptr()
    sentinel = 0xAB
    rec.terminal.fill_(sentinel)
    for _ in range(40):                                                # ~7 % of the envs end per step: plenty of terminal boards
        r = game.step(torch.randint(0, 4, (n,), generator=gen, device=game.device, dtype=torch.uint8))
        assert r.terminal_boards is None
    assert game._args.terminal_boards is None
    torch.cuda.synchronize()
    assert bool((rec.terminal == sentinel).all()), "a plain step() wrote into the recorder's terminal slice"
