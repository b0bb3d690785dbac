"""Draw-stream spec: Philox4x32-10 known answers and the spawn arithmetic (CPU only)."""
import numpy as np

from oracle import draws, oracle

# Random123 kat_vectors: philox4x32 10 rounds (ctr, key) -> output
KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_kat_numpy_and_c():
    for ctr, key, want in KATS:
        got = draws.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want
        got_c = oracle.philox([ctr], key[0], key[1])[0]
        assert tuple(int(x) for x in got_c) == want


def test_numpy_and_c_draw_words_agree():
    rng = np.random.default_rng(7)
    ctr = rng.integers(0, 2**32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    a = draws.philox4x32_10(ctr, (123, 456))
    b = oracle.philox(ctr, 123, 456)
    assert np.array_equal(a, b)


def test_counter_layout():
    env = np.array([0, 1, (1 << 32) + 5, (1 << 62) + 9], dtype=np.uint64)
    w = draws.draw_words(seed=(7 << 32) | 3, env_ids=env, idx=(2 << 32) | 11, tag=1)
    for i, e in enumerate(env):
        e = int(e)
        ctr = (11, 2, e & 0xFFFFFFFF, ((e >> 32) & 0x7FFFFFFF) | (1 << 31))
        assert np.array_equal(w[i], draws.philox4x32_10(np.array(ctr, dtype=np.uint32), (3, 7)))


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
