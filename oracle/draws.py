"""Draw-stream specification in numpy + draw injection into the unmodified reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* `philox4x32_10` / `philox2x32_10` / `draw_words` restate the draw stream frozen in
  include/g2048.h (version 2: a launch-uniform key from Philox4x32-10, one Philox2x32-10
  block per board; Salmon et al. SC'11, Random123 constants) with vectorised uint64
  arithmetic; pinned by the Random123 known-answer vectors in tests/test_draws.py.
* `InjectedDraws` is assigned to `env.np_random` of the UNMODIFIED reference env.
  The reference's add_tile (game2048_env.py:166-176) only calls `.random()` (:168)
  and `.shuffle(positions)` (:170); feeding it the kernel's draw word therefore makes
  the reference place exactly the tile the kernel places, while every other part of
  step() — move/shift/isend/reward/highest/stack — remains the reference's own code.
"""
import numpy as np

P2_THRESHOLD = 3865470567  # f < this  <=>  f / 2**32 < 0.9  (game2048_env.py:168)
M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (2,) ints -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = int(key[0]) & MASK32, int(key[1]) & MASK32
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & np.uint64(MASK32)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & np.uint64(MASK32)
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


M2 = 0xD256D193
TAG_STEP, TAG_RESET, TAG_POLICY = 0, 1, 2


def philox2x32_10(ctr, key):
    """ctr: (..., 2) uint32, key: (...) or scalar uint32 -> (..., 2) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c0, c1 = ctr[..., 0].copy(), ctr[..., 1].copy()
    key = np.broadcast_to(np.asarray(key, dtype=np.uint64) & np.uint64(MASK32), c0.shape).copy()
    for _ in range(10):
        p = np.uint64(M2) * c0
        c0 = (p >> np.uint64(32)) ^ key ^ c1
        c1 = p & np.uint64(MASK32)
        key = (key + np.uint64(W0)) & np.uint64(MASK32)
    return np.stack([c0, c1], axis=-1).astype(np.uint32)


def stream_keys(seed, env_ids, idx, tag):
    """The launch-uniform 32-bit key of every env id (it only depends on the id's high half)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    kctr = np.zeros(env_ids.shape + (4,), dtype=np.uint64)
    kctr[..., 0] = (int(idx) >> 32) & MASK32
    kctr[..., 1] = env_ids >> np.uint64(32)
    kctr[..., 2] = int(tag)
    return philox4x32_10(kctr, (int(seed) & MASK32, (int(seed) >> 32) & MASK32))[..., 0]


def draw_words(seed, env_ids, idx, tag):
    """Words w[0..3] for global env ids `env_ids` at step/reset index `idx` under stream `tag`."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    ctr = np.empty(env_ids.shape + (2,), dtype=np.uint64)
    ctr[..., 0] = env_ids & np.uint64(MASK32)
    ctr[..., 1] = int(idx) & MASK32
    x = philox2x32_10(ctr, stream_keys(seed, env_ids, idx, tag)).astype(np.uint64)
    w = np.zeros(env_ids.shape + (4,), dtype=np.uint64)
    w[..., 0] = x[..., 0]
    w[..., 1] = x[..., 1]
    w[..., 2] = (x[..., 1] << np.uint64(16)) & np.uint64(MASK32)
    return w.astype(np.uint32)


def spawn_from_word(n_empty, w):
    """(k, value) the draw word `w` selects on a board with `n_empty` empty cells."""
    p = int(w) * int(n_empty)
    k, f = p >> 32, p & MASK32
    return k, (2 if f < P2_THRESHOLD else 4)


def word_for_spawn(n_empty, k, value):
    """A draw word that makes spawn() put `value` on the k-th empty cell (CSV fixtures)."""
    n = int(n_empty)
    target = (k << 32) + (0 if value == 2 else P2_THRESHOLD)
    w = -(-target // n)               # smallest w with w*n >= target
    if w <= MASK32 and spawn_from_word(n, w) == (k, value):
        return w
    raise ValueError((n_empty, k, value))


class InjectedDraws:
    """Stand-in for `env.np_random` that replays one queue of draw words.

    Each add_tile() consumes exactly one word: `.random()` returns f / 2**32 so that
    the reference's `< 0.9` test (:168) picks the kernel's tile, and `.shuffle()`
    moves the k-th empty cell (row-major over env.Matrix) to the front of the
    position list so that the reference's first-empty scan (:171-175) lands on it.
    """

    def __init__(self, env):
        self.env = env
        self.words = []
        self._k = None

    def push(self, *words):
        self.words.extend(int(w) for w in words)

    def random(self):
        w = self.words.pop(0)
        n_empty = int((self.env.Matrix == 0).sum())
        p = w * n_empty
        self._k = p >> 32
        return (p & MASK32) / 4294967296.0

    def shuffle(self, positions):
        empties = [(r, c) for r in range(4) for c in range(4) if self.env.Matrix[r, c] == 0]
        target = empties[self._k]
        positions.remove(target)
        positions.insert(0, target)
        self._k = None
