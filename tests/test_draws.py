"""Draw-stream spec: Philox4x32-10 / Philox2x32-10 known answers and the spawn arithmetic (CPU only)."""
import numpy as np

from oracle import draws, oracle

# Random123 kat_vectors: philox4x32 10 rounds (ctr, key) -> output
KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
# Random123 kat_vectors: philox2x32 10 rounds (ctr, key) -> output
KATS_2X32 = [
    ((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
    ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
    ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12)),
]


def test_philox_kat_numpy_and_c():
    for ctr, key, want in KATS:
        got = draws.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want
        got_c = oracle.philox([ctr], key[0], key[1])[0]
        assert tuple(int(x) for x in got_c) == want
    for ctr, key, want in KATS_2X32:
        got = draws.philox2x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want
        got_c = oracle.philox2x32([ctr], key)[0]
        assert tuple(int(x) for x in got_c) == want


def test_numpy_and_c_draw_words_agree():
    rng = np.random.default_rng(7)
    ctr = rng.integers(0, 2**32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    a = draws.philox4x32_10(ctr, (123, 456))
    b = oracle.philox(ctr, 123, 456)
    assert np.array_equal(a, b)
    ctr2 = rng.integers(0, 2**32, size=(4096, 2), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(draws.philox2x32_10(ctr2, 77), oracle.philox2x32(ctr2, 77))
    import parity_checks as pc
    from backends import OracleOps
    pc.check_draw_words(OracleOps)


def test_counter_layout():
    """include/g2048.h "Draw stream", written out longhand for a few ids."""
    env = np.array([0, 1, (1 << 32) + 5, (1 << 62) + 9], dtype=np.uint64)
    seed, idx = (7 << 32) | 3, (2 << 32) | 11
    for tag in (0, 1, 2):
        w = draws.draw_words(seed=seed, env_ids=env, idx=idx, tag=tag)
        for i, e in enumerate(env):
            e = int(e)
            key = draws.philox4x32_10(np.array((2, e >> 32, tag, 0), dtype=np.uint32), (3, 7))[0]
            x = draws.philox2x32_10(np.array((e & 0xFFFFFFFF, 11), dtype=np.uint32), int(key))
            assert [int(v) for v in w[i]] == [int(x[0]), int(x[1]), (int(x[1]) << 16) & 0xFFFFFFFF, 0]
    # streams of different tags, seeds, indices and id halves are unrelated
    a = draws.draw_words(seed, env, idx, 0)
    for other in (draws.draw_words(seed, env, idx, 1), draws.draw_words(seed + 1, env, idx, 0),
                  draws.draw_words(seed, env, idx + (1 << 32), 0), draws.draw_words(seed, env + np.uint64(1 << 32), idx, 0)):
        assert not np.any(a[:, :2] == other[:, :2])


def test_reset_words_are_independent_enough():
    """w[1] = x1 and w[2] = x1 << 16 drive the two spawns of a reset: cell and tile of the second
    spawn must be uniform / P(4)=0.1 whatever the first spawn drew."""
    w = draws.draw_words(3, np.arange(400000, dtype=np.uint64), 9, 1).astype(np.uint64)
    k1 = (w[:, 1] * np.uint64(16)) >> np.uint64(32)
    four1 = ((w[:, 1] * np.uint64(16)) & np.uint64(0xFFFFFFFF)) >= np.uint64(draws.P2_THRESHOLD)
    k2 = (w[:, 2] * np.uint64(15)) >> np.uint64(32)
    four2 = ((w[:, 2] * np.uint64(15)) & np.uint64(0xFFFFFFFF)) >= np.uint64(draws.P2_THRESHOLD)
    n = len(w)
    assert abs(four1.mean() - 0.1) < 0.002 and abs(four2.mean() - 0.1) < 0.002
    joint = np.zeros((16, 15))
    np.add.at(joint, (k1.astype(int), k2.astype(int)), 1)
    assert np.abs(joint / n * 240 - 1).max() < 0.15            # every (cell, cell) pair equally likely
    assert abs(four2[four1].mean() - 0.1) < 0.006 and abs(four2[~four1].mean() - 0.1) < 0.002
    for k in range(15):
        assert abs(four1[k2 == k].mean() - 0.1) < 0.01


def test_p2_threshold_is_reference_comparison():
    # random() < 0.9 (game2048_env.py:168) with random() = f / 2**32
    t = draws.P2_THRESHOLD
    assert (t - 1) / 4294967296.0 < 0.9 and not (t / 4294967296.0 < 0.9)


def test_spawn_word_roundtrip_and_uniformity():
    for n in range(1, 17):
        for k in range(n):
            for v in (2, 4):
                assert draws.spawn_from_word(n, draws.word_for_spawn(n, k, v)) == (k, v)
    rng = np.random.default_rng(0)
    w = rng.integers(0, 2**32, size=200000, dtype=np.uint64)
    p = w * np.uint64(7)
    k = (p >> np.uint64(32)).astype(np.int64)
    f = (p & np.uint64(0xFFFFFFFF))
    assert k.min() == 0 and k.max() == 6
    assert abs(np.bincount(k).std() / np.bincount(k).mean()) < 0.02
    assert abs((f >= draws.P2_THRESHOLD).mean() - 0.1) < 0.005
