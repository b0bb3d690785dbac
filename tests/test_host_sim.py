"""The byte-SIMD device code (gym-2048_b200/csrc/g2048_device.cuh), compiled for the host
with prmt/umulhi/popc emulated, against the golden vectors and the oracle.  This is the
GPU-less check of the exact logic the CUDA kernels execute; tests/test_gpu_parity.py runs
the same checks on the device through the C ABI."""
import parity_checks as pc
from backends import SimOps


def test_philox():
    pc.check_philox(SimOps)


def test_shift_table_all_directions():
    pc.check_shift_table_all_directions(SimOps)


def test_csv_transitions():
    pc.check_csv_transitions(SimOps)


def test_special_boards():
    pc.check_special_boards(SimOps)


def test_golden_rollouts(golden_rollouts):
    pc.check_rollouts(SimOps, golden_rollouts)


def test_status_and_move_random_boards():
    pc.check_status_random_boards(SimOps)


def test_add_tile():
    pc.check_add_tile(SimOps)


def test_rollout_vs_oracle_random():
    pc.check_against_oracle(SimOps, n=4096, steps=48, seed=0, policy="random")


def test_rollout_vs_oracle_legal_midgame():
    pc.check_against_oracle(SimOps, n=2048, steps=400, seed=42, policy="legal", illegal_move_reward=-1.0,
                            env_id_base=(1 << 40) + 3)


def test_rollout_vs_oracle_max_tile_noreset():
    pc.check_against_oracle(SimOps, n=1024, steps=300, seed=456, policy="legal", max_tile_exp=6,
                            auto_reset=False)


def test_symmetries_match_reference_and_oracle():
    """training_data.hflip / rotate PRMT networks vs the reference fixtures and the oracle."""
    import ctypes as C
    import numpy as np
    from backends import sim_lib, _p
    from conftest import load_golden
    from oracle import oracle
    gold = load_golden("transitions.npz")
    x, y = gold["syn/in_x"], gold["syn/in_y"]
    for tag, h, k in (("hflip", 1, 0), ("rot1", 0, 1), ("rot2", 0, 2), ("rot3", 0, 3), ("hflip_rot3", 1, 3)):
        ob, oa = np.empty_like(x), np.empty_like(y)
        sim_lib().sim_symmetry(_p(x), _p(ob), _p(y), _p(oa), C.c_uint64(len(x)), C.c_int(h), C.c_int(k))
        assert np.array_equal(ob, gold["syn/%s_x" % tag]) and np.array_equal(oa, gold["syn/%s_y" % tag]), tag
    rng = np.random.default_rng(5)
    b = rng.integers(0, 19, (5000, 16)).astype(np.uint8)
    a = rng.integers(0, 4, 5000).astype(np.uint8)
    for h in (0, 1):
        for k in range(4):
            ob, oa = np.empty_like(b), np.empty_like(a)
            sim_lib().sim_symmetry(_p(b), _p(ob), _p(a), _p(oa), C.c_uint64(len(b)), C.c_int(h), C.c_int(k))
            rb, ra = oracle.symmetry(b, a, h, k)
            assert np.array_equal(ob, rb) and np.array_equal(oa, ra), (h, k)


def test_sample_actions_matches_oracle():
    import ctypes as C
    import numpy as np
    from backends import sim_lib, _p
    from oracle import oracle
    n = 50000
    masks = np.random.default_rng(9).integers(0, 16, n).astype(np.uint8)
    for m in (masks, None):
        out = np.zeros(n, np.uint8)
        sim_lib().sim_sample_actions(_p(m), _p(out), C.c_uint64(n), C.c_uint64(123), C.c_uint64(42), C.c_uint64(17))
        assert np.array_equal(out, oracle.sample_actions(m, n, 123, 42, 17))


def test_draw_words_all_forms():
    """draw_words, DrawStream (cached key) and the step kernel's precomputed round keys give the
    words of the spec."""
    import parity_checks as pc

    for form in (0, 1, 2):
        class Ops:
            draw_words = staticmethod(lambda n, base, seed, idx, tag, form=form: SimOps.draw_words(n, base, seed, idx, tag, form))
        pc.check_draw_words(Ops)


def test_fresh_board_table_equals_two_spawns():
    """The 1024-entry fresh-board table of the step kernels (two_tile_board, also evaluated at compile time for
    the device constant) gives the board the two reset spawns give, for every cell pair and tile pair."""
    import ctypes as C
    import numpy as np
    from backends import sim_lib, _p
    rng = np.random.default_rng(3)
    n = 400000
    w1 = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    w2 = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    w1[:8] = [0, 0xFFFFFFFF, 0xF0000000, 0x0FFFFFFF, 0xE6666666, 0xE6666667, 0x10000000, 0x80000000]
    w2[:8] = [0xFFFFFFFF, 0, 0xFFFFFFFF, 0x11111111, 0xFFFFFFFF, 0, 0xEEEEEEEE, 0x77777777]
    a, b = np.zeros((n, 16), np.uint8), np.zeros((n, 16), np.uint8)
    sim_lib().sim_fresh_boards(_p(w1), _p(w2), _p(a), _p(b), C.c_uint64(n))
    assert np.array_equal(a, b)
    assert np.all((a != 0).sum(axis=1) == 2) and set(np.unique(a)) == {0, 1, 2}
    seen = {(tuple(np.flatnonzero(r)), tuple(r[r != 0])) for r in a[:200000]}
    assert len(seen) == 16 * 15 // 2 * 4                        # every cell pair with every tile pair occurs

