// g2048_device.cuh — device-side building blocks of the batched 2048 step (sm_100a).
//
// A board is four 32-bit registers r0..r3 (row i = r[i], byte j = cell (i,j)), each
// byte the tile EXPONENT (0 = empty, value 2^e).  All game logic is byte-SIMD inside
// 32-bit registers: the four lines of a move are processed at once, one line per
// byte lane.  Exponents never exceed 0x3F, so bit 7 of every byte is free and is
// used as the per-byte carry/flag bit:
//     x + 0x7F7F7F7F   has bit 7 of byte j set   <=>   byte j of x is non-zero
//     prmt(x, 0xBA98)  replicates bit 7 of each byte over that byte (0x00 / 0xFF)
// No byte-wise op below can carry into the neighbouring byte (each proof is next to
// the op).  Reference semantics: /root/reference/env/envs/game2048_env.py (file:line
// cited per function); none of that code is used — the reference walks Python lists.
#pragma once
#include <cstdint>

// tests/host_sim compiles this header with g++ (G2048_HOST_SIM) to unit-test the exact
// byte-SIMD logic on the CPU box; the product only ever compiles it with nvcc.
#ifdef G2048_HOST_SIM
#define G2048_DEV inline
#define G2048_CONST static const
#else
#define G2048_DEV __device__ __forceinline__
#define G2048_CONST __constant__
#endif

namespace g2048 {

constexpr uint32_t H = 0x80808080u;   // bit 7 of every byte
constexpr uint32_t L7 = 0x7F7F7F7Fu;  // low 7 bits of every byte
constexpr uint32_t K1 = 0x01010101u;  // 1 in every byte
constexpr uint32_t P2_THRESHOLD = 3865470567u;  // f < T  <=>  f/2^32 < 0.9 (:168)

#ifdef G2048_HOST_SIM
// PTX prmt.b32 default mode: nibble i of sel picks byte (sel & 7) of {b,a}; bit 3 of the
// nibble replicates that byte's sign bit instead.
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t d = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t c = (sel >> (4 * i)) & 0xF;
    uint32_t byte = (uint32_t)(src >> (8 * (c & 7))) & 0xFF;
    if (c & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
    d |= byte << (8 * i);
  }
  return d;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t __popc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
#else
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
#endif
// 0xFF in every byte whose bit 7 is set, else 0x00.
G2048_DEV uint32_t spread(uint32_t x) { return prmt(x, 0u, 0xBA98u); }
// 0xFF where the byte of x is non-zero (bytes of x <= 0x80).
G2048_DEV uint32_t nzmask(uint32_t x) { return spread(x + L7); }

// ---- Philox4x32-10 (Salmon et al. SC'11) ------------------------------------------
struct Words { uint32_t w0, w1, w2, w3; };

G2048_DEV Words philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;   // key schedule is warp-uniform: folded into constants
    k1 += 0xBB67AE85u;
  }
  return Words{c0, c1, c2, c3};
}

// Draw words for (seed, global env id, index, tag) — include/g2048.h "Draw stream".
G2048_DEV Words draw_words(uint64_t seed, uint64_t env_id, uint64_t idx, uint32_t tag) {
  return philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)env_id,
                       ((uint32_t)(env_id >> 32) & 0x7FFFFFFFu) | (tag << 31),
                       (uint32_t)seed, (uint32_t)(seed >> 32));
}

// ---- orientation --------------------------------------------------------------------
// move() (:210-237) shifts every line toward one side.  We rotate the board into a
// frame (a,b,c,d) where cell words are ordered along the move (a = destination end)
// and the four lines sit in the four byte lanes, using a two-stage PRMT network whose
// selectors depend on the action.  Stage 1 pairs (x,z) and (y,w); stage 2 pairs the
// results.  Up/Down are word permutations, Left/Right a byte transpose (Right with the
// bytes reversed first).  kOrientIn[action] = {A, B, C, D}:
//   x0 = prmt(r0,r2,A) x1 = prmt(r0,r2,B) y0 = prmt(r1,r3,A) y1 = prmt(r1,r3,B)
//   a  = prmt(x0,y0,C) b  = prmt(x0,y0,D) c  = prmt(x1,y1,C) d  = prmt(x1,y1,D)
// The inverse network has the same shape (kOrientOut).  Action map (:196, :210-212):
// 0 = Up (toward row 0), 1 = Right (col 3), 2 = Down (row 3), 3 = Left (col 0).
struct alignas(16) Sel4 { uint32_t A, B, C, D; };
G2048_CONST Sel4 kOrientIn[4] = {
    {0x3210u, 0x7654u, 0x3210u, 0x7654u},   // Up:    identity
    {0x6273u, 0x4051u, 0x5140u, 0x7362u},   // Right: reverse bytes, transpose
    {0x7654u, 0x3210u, 0x7654u, 0x3210u},   // Down:  reverse words
    {0x5140u, 0x7362u, 0x5140u, 0x7362u},   // Left:  transpose
};
G2048_CONST Sel4 kOrientOut[4] = {
    {0x3210u, 0x7654u, 0x3210u, 0x7654u},
    {0x1504u, 0x3726u, 0x1504u, 0x3726u},   // transpose, reverse bytes
    {0x7654u, 0x3210u, 0x7654u, 0x3210u},
    {0x5140u, 0x7362u, 0x5140u, 0x7362u},
};

G2048_DEV void orient(const Sel4 s, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                       uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  const uint32_t x0 = prmt(r0, r2, s.A), x1 = prmt(r0, r2, s.B);
  const uint32_t y0 = prmt(r1, r3, s.A), y1 = prmt(r1, r3, s.B);
  a = prmt(x0, y0, s.C);
  b = prmt(x0, y0, s.D);
  c = prmt(x1, y1, s.C);
  d = prmt(x1, y1, s.D);
}

// ---- shift (:243-260) on four lines at once -----------------------------------------
// if x == 0 (per byte): x <- y, y <- 0
G2048_DEV void bubble(uint32_t& x, uint32_t& y) {
  const uint32_t m = nzmask(x);
  x |= y & ~m;
  y &= m;
}

// Slide (a,b,c,d) toward a with merging; returns the move score (:254).
G2048_DEV uint32_t slide_merge(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  // compaction: bubble the zeros toward d (3+2+1 conditional moves) (:250-251)
  bubble(a, b); bubble(b, c); bubble(c, d);
  bubble(a, b); bubble(b, c);
  bubble(a, b);
  // merges on the compacted line: leftmost pair first, each tile merges once (:252-259)
  //   m1: a==b!=0;  m2: b==c!=0 and not m1;  m3: c==d!=0 and not m2
  const uint32_t m1 = ~((a ^ b) + L7) & (b + L7);          // bit 7 only is meaningful
  const uint32_t m2 = ~((b ^ c) + L7) & (c + L7) & ~m1;
  const uint32_t m3 = ~((c ^ d) + L7) & (d + L7) & ~m2;
  const uint32_t M1 = spread(m1), M2 = spread(m2), M3 = spread(m3);
  a += M1 & K1;                                            // exponent + 1 (value * 2, :253)
  b = (b & ~M1) + (M2 & K1);
  c = (c & ~M2) + (M3 & K1);
  d &= ~M3;
  // merged exponents: per lane either (m1 and maybe m3) or m2 alone -> two slot words
  const uint32_t sA = (a & M1) | (b & M2);
  const uint32_t sB = c & M3;
  // close the holes the merges left at b and/or c, d
  {
    const uint32_t m = nzmask(b);
    b |= c & ~m;
    c = (c & m) | (d & ~m);
    d &= m;
  }
  bubble(c, d);
  // score = sum over the 8 slots of 2^e (e != 0).  Shifts use e mod 32 (e <= 18).
  uint32_t s = (1u << (sA & 31)) + (1u << ((sA >> 8) & 31)) + (1u << ((sA >> 16) & 31)) +
               (1u << ((sA >> 24) & 31)) + (1u << (sB & 31)) + (1u << ((sB >> 8) & 31)) +
               (1u << ((sB >> 16) & 31)) + (1u << ((sB >> 24) & 31));
  // empty slots contributed 2^0 each: subtract them
  const uint32_t filled = __popc(((M1 | M2) & K1) | (M3 & 0x02020202u));
  return s - 8u + filled;
}

// ---- add_tile (:166-176) under the draw-stream definition ----------------------------
// Places 2 (P=0.9) or 4 on the k-th empty cell (row-major), k = (w * n_empty) >> 32.
// Returns n_empty BEFORE the spawn (0 => nothing placed).
// `enable` = 0xFFFFFFFF to place the tile, 0 to leave the board untouched (illegal move:
// no tile, :91-95).
G2048_DEV uint32_t spawn(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t w,
                         uint32_t enable = 0xFFFFFFFFu) {
  // e_i: bit 7 set where the cell is empty; q_i: 0/1 per byte
  const uint32_t e0 = ~(r0 + L7) & H, e1 = ~(r1 + L7) & H, e2 = ~(r2 + L7) & H, e3 = ~(r3 + L7) & H;
  const uint32_t q0 = e0 >> 7, q1 = e1 >> 7, q2 = e2 >> 7, q3 = e3 >> 7;
  // inclusive row-major prefix counts of empties, one per byte (<= 16: no carries);
  // byte 3 of each word is the running total, broadcast into the next row's offset
  const uint32_t p0 = q0 * K1;
  const uint32_t p1 = q1 * K1 + prmt(p0, 0u, 0x3333u);
  const uint32_t p2 = q2 * K1 + prmt(p1, 0u, 0x3333u);
  const uint32_t p3 = q3 * K1 + prmt(p2, 0u, 0x3333u);
  const uint32_t n = p3 >> 24;
  const uint32_t k = __umulhi(w, n);
  const uint32_t f = w * n;
  // target = the cell with prefix == k+1 that is empty:  (p > k) and not (p > k+1)
  const uint32_t gk = L7 - k * K1;          // p + gk has bit 7  <=>  p >= k+1
  const uint32_t gk1 = gk - K1;             // p + gk1 has bit 7 <=>  p >= k+2
  const uint32_t tile = ((f < P2_THRESHOLD) ? K1 : 0x02020202u) & enable;   // exponent 1 or 2 (:168)
  r0 |= spread((p0 + gk) & ~(p0 + gk1) & e0) & tile;
  r1 |= spread((p1 + gk) & ~(p1 + gk1) & e1) & tile;
  r2 |= spread((p2 + gk) & ~(p2 + gk1) & e2) & tile;
  r3 |= spread((p3 + gk) & ~(p3 + gk1) & e3) & tile;
  return n;
}

// reset (:102-111): zero board, two spawns from w1, w2.  Specialised: the first spawn
// sees 16 empties (k1 = w1 >> 28), the second 15 and skips cell k1.
G2048_DEV void fresh_board(uint32_t w1, uint32_t w2, uint32_t& r0, uint32_t& r1,
                                            uint32_t& r2, uint32_t& r3) {
  const uint32_t k1 = w1 >> 28;
  const uint32_t t1 = ((w1 << 4) < P2_THRESHOLD) ? 1u : 2u;
  uint32_t k2 = __umulhi(w2, 15u);
  const uint32_t t2 = ((w2 * 15u) < P2_THRESHOLD) ? 1u : 2u;
  k2 += (k2 >= k1) ? 1u : 0u;
  // 128-bit one-hot byte insert done as two 64-bit halves
  const uint64_t v1 = (uint64_t)t1 << ((k1 & 7u) * 8u), v2 = (uint64_t)t2 << ((k2 & 7u) * 8u);
  const uint64_t lo = ((k1 < 8u) ? v1 : 0ull) | ((k2 < 8u) ? v2 : 0ull);
  const uint64_t hi = ((k1 < 8u) ? 0ull : v1) | ((k2 < 8u) ? 0ull : v2);
  r0 = (uint32_t)lo; r1 = (uint32_t)(lo >> 32); r2 = (uint32_t)hi; r3 = (uint32_t)(hi >> 32);
}

// ---- board queries --------------------------------------------------------------------
// highest (:190-192) as exponent.  max per byte: a + 0x80 - b keeps every byte in
// [0x80-63, 0x80+63] (no carry/borrow) and has bit 7 set <=> a >= b.
G2048_DEV uint32_t bmax(uint32_t a, uint32_t b) {
  const uint32_t m = spread(a + H - b);
  return (a & m) | (b & ~m);
}
G2048_DEV uint32_t highest_exp(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  uint32_t m = bmax(bmax(r0, r1), bmax(r2, r3));
  m = bmax(m, m >> 16);
  m = bmax(m, m >> 8);
  return m & 0xFFu;
}

// "no legal move" on a board WITHOUT empty cells: no two adjacent cells are equal.
// (isend :270-280: with an empty cell it returns False before trying moves; on a full
// board a move is legal iff two neighbours along it are equal.)
G2048_DEV bool full_board_is_dead(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  // x + L7 has bit 7 clear exactly where the byte of x is zero (equal neighbours).
  // Horizontal: byte j of r ^ (r >> 8) compares cells j and j+1; byte 3 is the cell
  // itself, non-zero on a full board.
  const uint32_t v = ((r0 ^ r1) + L7) & ((r1 ^ r2) + L7) & ((r2 ^ r3) + L7);
  const uint32_t h = ((r0 ^ (r0 >> 8)) + L7) & ((r1 ^ (r1 >> 8)) + L7) & ((r2 ^ (r2 >> 8)) + L7) &
                     ((r3 ^ (r3 >> 8)) + L7);
  return ((~(v & h)) & H) == 0u;
}

// Legal-move mask: bit d set <=> move(d, trial=True) does not raise (:224,236-239).
// A line moves toward its head iff some cell is empty with a tile right behind it, or
// two adjacent tiles are equal.
G2048_DEV uint32_t legal_mask(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  const uint32_t n0 = r0 + L7, n1 = r1 + L7, n2 = r2 + L7, n3 = r3 + L7;   // bit7: non-empty
  // vertical pairs (upper u, lower l)
  const uint32_t eqv = (~((r0 ^ r1) + L7) & n0) | (~((r1 ^ r2) + L7) & n1) | (~((r2 ^ r3) + L7) & n2);
  const uint32_t up = (~n0 & n1) | (~n1 & n2) | (~n2 & n3) | eqv;     // hole above a tile
  const uint32_t dn = (n0 & ~n1) | (n1 & ~n2) | (n2 & ~n3) | eqv;     // hole below a tile
  // horizontal pairs: byte j of s_i is cell (i, j+1); only byte lanes 0..2 are pairs
  const uint32_t s0 = r0 >> 8, s1 = r1 >> 8, s2 = r2 >> 8, s3 = r3 >> 8;
  const uint32_t m0 = s0 + L7, m1 = s1 + L7, m2 = s2 + L7, m3 = s3 + L7;
  const uint32_t eqh = (~((r0 ^ s0) + L7) & n0) | (~((r1 ^ s1) + L7) & n1) |
                       (~((r2 ^ s2) + L7) & n2) | (~((r3 ^ s3) + L7) & n3);
  const uint32_t lf = (~n0 & m0) | (~n1 & m1) | (~n2 & m2) | (~n3 & m3) | eqh;   // hole left of a tile
  const uint32_t rt = (n0 & ~m0) | (n1 & ~m1) | (n2 & ~m2) | (n3 & ~m3) | eqh;   // hole right of a tile
  constexpr uint32_t HP = 0x00808080u;                                           // pair lanes only
  return ((up & H) ? 1u : 0u) | ((rt & HP) ? 2u : 0u) | ((dn & H) ? 4u : 0u) | ((lf & HP) ? 8u : 0u);
}

G2048_DEV uint32_t count_empty(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  return 16u - __popc((r0 + L7) & H) - __popc((r1 + L7) & H) - __popc((r2 + L7) & H) - __popc((r3 + L7) & H);
}

// ---- one whole step (:76-100) on a board held in registers --------------------------
struct StepOut {
  uint32_t score;     // merge score of the move (0 when illegal)
  uint32_t highest;   // exponent of the highest tile after the spawn (valid if requested)
  bool legal;         // False <=> IllegalMove (:91-95)
  bool done;          // terminated
  uint32_t t0, t1, t2, t3;   // post-spawn board (the terminal board when done)
};

// On return r0..r3 hold the board handed back to the agent: the post-spawn board, or a
// fresh reset() board when the episode ended and auto_reset is set (SB3 DummyVecEnv).
G2048_DEV StepOut step_board(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t action,
                             const Words& w, uint32_t max_tile_exp, bool want_highest, bool auto_reset) {
  StepOut o;
  uint32_t a, b, c, d;
  orient(kOrientIn[action], r0, r1, r2, r3, a, b, c, d);
  const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
  o.score = slide_merge(a, b, c, d);
  // :238-239 — nothing changed => IllegalMove.  An illegal move leaves (a,b,c,d) as they
  // were and scores 0, so the same data path serves both cases; only the spawn is gated.
  o.legal = (((a ^ a0) | (b ^ b0)) | ((c ^ c0) | (d ^ d0))) != 0u;
  orient(kOrientOut[action], a, b, c, d, r0, r1, r2, r3);
  const uint32_t n_empty = spawn(r0, r1, r2, r3, w.w0, o.legal ? 0xFFFFFFFFu : 0u);     // :88
  o.highest = 0;
  if (want_highest || max_tile_exp != 0u) o.highest = highest_exp(r0, r1, r2, r3);      // :97
  // isend (:262-280) for a legal move; an illegal move terminates (:94)
  bool end = (n_empty == 1u) && full_board_is_dead(r0, r1, r2, r3);
  if (max_tile_exp != 0u) end = end || (o.highest == max_tile_exp);                      // :267
  o.done = end || !o.legal;
  o.t0 = r0; o.t1 = r1; o.t2 = r2; o.t3 = r3;
  if (auto_reset && o.done) fresh_board(w.w1, w.w2, r0, r1, r2, r3);                    // :102-111
  return o;
}

}  // namespace g2048
