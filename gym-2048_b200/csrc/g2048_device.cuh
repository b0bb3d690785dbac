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
#define G2048_HD inline
#define G2048_CONST static const
#define G2048_HD_CONSTEXPR constexpr inline
#else
#define G2048_DEV __device__ __forceinline__
#define G2048_HD __host__ __device__ __forceinline__
#define G2048_CONST static __constant__      // one copy per translation unit (no relocatable device code)
#define G2048_HD_CONSTEXPR __host__ __device__ constexpr      // also evaluated at compile time (the fresh-board table)
#endif

// EXPERIMENTS ONLY (scripts/kernel_variants.py): a bit mask that REMOVES parts of the step so that their marginal
// cost can be measured in place.  Any non-zero value computes wrong results; the product is built with 0.
//   1 Philox without wide multiplies   2 no tile insertion   8 no score   16 no Philox   32 no spawn
//   64 no merge   128 no compaction + merge   256 no orientation networks
#ifndef G2048_ABLATE
#define G2048_ABLATE 0
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
inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) {         // shf.r.clamp: shift count min(sh, 32)
  return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh > 32u ? 32u : sh));
}
#else
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
#endif
// Pipe balancing.  The step is bound by the integer ALU pipe (LOP3/PRMT/SHF/IADD3 issue at
// half rate); the FMA pipe, where IMAD runs, is mostly idle.  addf() is an add that ptxas
// must keep as IMAD (x * one + y with `one` read from constant memory, so it cannot be
// strength-reduced back to IADD3), shr() a right shift done as IMAD.HI.
#ifdef G2048_HOST_SIM
inline uint32_t addf(uint32_t x, uint32_t y) { return x + y; }
inline uint32_t addm(uint32_t x, uint32_t y) { return x + y; }
template <int S> inline uint32_t shr(uint32_t x) { return x >> S; }
template <int S> inline uint32_t shl(uint32_t x) { return x << S; }
inline uint32_t madhi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)(((uint64_t)a * b) >> 32) + c; }
#else
static __constant__ uint32_t kOne = 1u;
#ifndef G2048_ADDF_PLAIN      // 1: addf() is a plain add (ptxas picks IADD3 / IMAD itself) — experiment
#define G2048_ADDF_PLAIN 1
#endif
__device__ __forceinline__ uint32_t addf(uint32_t x, uint32_t y) {
#if G2048_ADDF_PLAIN
  return x + y;
#else
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(kOne), "r"(y));
  return d;
#endif
}
#ifndef G2048_SHR_ALU         // 1: constant right shifts as SHF (ALU pipe) instead of IMAD.HI (FMA-heavy, quarter rate)
#define G2048_SHR_ALU 1
#endif
// An add that stays IMAD (FMA-heavy pipe) whatever ptxas would choose: for the legal-move mask, whose 45 extra
// instructions are two thirds LOP3 — with its 19 adds left to ptxas (IADD3 on either pipe) the mask kernels are 6 %
// slower (1 Mi boards, random-legal policy: 14.7 vs 13.8 us chained, 16.3 vs 15.4 us plain; profiles/r02_variants.log).
#ifndef G2048_MASK_ADD_FMA
#define G2048_MASK_ADD_FMA 1
#endif
__device__ __forceinline__ uint32_t addm(uint32_t x, uint32_t y) {
#if G2048_MASK_ADD_FMA
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(kOne), "r"(y));
  return d;
#else
  return x + y;
#endif
}
template <int S> __device__ __forceinline__ uint32_t shr(uint32_t x) {
  return G2048_SHR_ALU ? (x >> S) : __umulhi(x, 1u << (32 - S));
}
template <int S> __device__ __forceinline__ uint32_t shl(uint32_t x) { return x * (1u << S); }
// hi32(a * b) + c in one IMAD.HI
__device__ __forceinline__ uint32_t madhi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
#endif
// 0xFF in every byte whose bit 7 is set, else 0x00.
G2048_DEV uint32_t spread(uint32_t x) { return prmt(x, 0u, 0xBA98u); }
// 0xFF where the byte of x is non-zero (bytes of x <= 0x80).
G2048_DEV uint32_t nzmask(uint32_t x) { return spread(addf(x, L7)); }

// ---- draw stream: Philox (Salmon et al. SC'11, Random123 constants) ---------------------------
// Two generators of the family.  Philox4x32-10 runs ONCE PER LAUNCH (on the host, or in the
// kernel prologue when the step index lives on the device) and compresses everything that is
// the same for every board of a launch — seed, high halves of the step index and of the env id,
// stream tag — into one 32-bit key.  Philox2x32-10, keyed with it and counting (env_lo, idx_lo),
// runs once per board: 10 wide multiplies instead of 20, and 64 output bits are all a step needs
// (32 for the spawn, 32 for the two spawns of a reset).
struct Words { uint32_t w0, w1, w2, w3; };

G2048_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

G2048_HD Words philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Words{c0, c1, c2, c3};
}

constexpr uint32_t PHILOX2_M = 0xD256D193u, PHILOX_W = 0x9E3779B9u;
struct Pair { uint32_t x0, x1; };

G2048_HD Pair philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi = mulhi32(PHILOX2_M, c0), lo = PHILOX2_M * c0;
    c0 = hi ^ key ^ c1;
    c1 = lo;
    key += PHILOX_W;
  }
  return Pair{c0, c1};
}

// Stream tags: which consumer the words are for.
constexpr uint32_t TAG_STEP = 0u, TAG_RESET = 1u, TAG_POLICY = 2u;

// The launch-uniform key (include/g2048.h "Draw stream").
G2048_HD uint32_t stream_key(uint64_t seed, uint64_t idx, uint64_t env_id, uint32_t tag) {
  return philox4x32_10((uint32_t)(idx >> 32), (uint32_t)(env_id >> 32), tag, 0u, (uint32_t)seed,
                       (uint32_t)(seed >> 32)).w0;
}

// The per-board generator with the ten round keys precomputed (host side, passed in
// kernel-parameter constant memory, so the key schedule costs no instruction in the hot loop)
// and counter word 1 — the low half of the step index, the same for every board — folded into
// round 0's key: round 0 is then one wide multiply and one two-input XOR.
struct StreamKeys { uint32_t k[10]; };        // k[0] = key ^ idx_lo, k[r] = key + r * W
G2048_HD void make_stream_keys(uint32_t key, uint32_t idx_lo, StreamKeys& ks) {
  ks.k[0] = key ^ idx_lo;
  for (int r = 1; r < 10; ++r) ks.k[r] = key + (uint32_t)r * PHILOX_W;
}
// k0 = key ^ idx_lo given explicitly (a kernel that runs several step indices per launch keeps
// k[1..9] in constant memory and derives k0 per step).
G2048_DEV Pair philox2x32_10_keys(uint32_t env_lo, uint32_t k0, const StreamKeys& ks) {
  uint32_t c0 = mulhi32(PHILOX2_M, env_lo) ^ k0, c1 = PHILOX2_M * env_lo;
#pragma unroll
  for (int r = 1; r < 10; ++r) {
    const uint32_t hi = mulhi32(PHILOX2_M, c0), lo = PHILOX2_M * c0;
    c0 = hi ^ ks.k[r] ^ c1;
    c1 = lo;
  }
  return Pair{c0, c1};
}
G2048_DEV Pair philox2x32_10_keys(uint32_t env_lo, const StreamKeys& ks) {
#if G2048_ABLATE & 16
  return Pair{env_lo * 0x9E3779B9u ^ ks.k[0], env_lo * 0x85EBCA6Bu ^ ks.k[1]};
#elif G2048_ABLATE & 1
  uint32_t c0 = (0x85EBCA6Bu * env_lo) ^ ks.k[0], c1 = PHILOX2_M * env_lo;
#pragma unroll
  for (int r = 1; r < 10; ++r) {
    const uint32_t hi = 0x85EBCA6Bu * c0, lo = PHILOX2_M * c0;
    c0 = hi ^ ks.k[r] ^ c1;
    c1 = lo;
  }
  return Pair{c0, c1};
#else
  uint32_t c0 = mulhi32(PHILOX2_M, env_lo) ^ ks.k[0], c1 = PHILOX2_M * env_lo;
#pragma unroll
  for (int r = 1; r < 10; ++r) {
    const uint32_t hi = mulhi32(PHILOX2_M, c0), lo = PHILOX2_M * c0;
    c0 = hi ^ ks.k[r] ^ c1;
    c1 = lo;
  }
  return Pair{c0, c1};
#endif
}

// Draw words from the generator output: w0 = x0 spawns after a legal move; a reset spawns with
// w1 = x1 and then w2 = x1 << 16 (the first spawn reads x1 from the top — 4 bits of cell, then the
// 2-or-4 fraction —, the second starts at bit 15; the two only share bits below 2^-12 of the first
// fraction's resolution).  w3 is unused (0).
G2048_HD Words words_from_pair(const Pair x) { return Words{x.x0, x.x1, x.x1 << 16, 0u}; }

// Draw words for (seed, global env id, index, tag) — include/g2048.h "Draw stream".
G2048_HD Words draw_words(uint64_t seed, uint64_t env_id, uint64_t idx, uint32_t tag) {
  return words_from_pair(philox2x32_10((uint32_t)env_id, (uint32_t)idx, stream_key(seed, idx, env_id, tag)));
}

// The same for a loop over many boards of one (seed, idx, tag): the launch-uniform key is
// recomputed only when the high half of the env id changes (at most once per 2^32 boards).
struct DrawStream {
  uint64_t seed, idx;
  uint32_t tag, env_hi, key;
  bool have_key;
  G2048_HD DrawStream(uint64_t seed_, uint64_t idx_, uint32_t tag_)
      : seed(seed_), idx(idx_), tag(tag_), env_hi(0u), key(0u), have_key(false) {}
  G2048_HD Words words(uint64_t env_id) {
    const uint32_t hi = (uint32_t)(env_id >> 32);
    if (!have_key || hi != env_hi) {
      env_hi = hi;
      key = stream_key(seed, idx, env_id, tag);
      have_key = true;
    }
    return words_from_pair(philox2x32_10((uint32_t)env_id, (uint32_t)idx, key));
  }
};

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
#if G2048_ABLATE & 256
  a = r0 ^ s.A; b = r1; c = r2; d = r3;
#else
  const uint32_t x0 = prmt(r0, r2, s.A), x1 = prmt(r0, r2, s.B);
  const uint32_t y0 = prmt(r1, r3, s.A), y1 = prmt(r1, r3, s.B);
  a = prmt(x0, y0, s.C);
  b = prmt(x0, y0, s.D);
  c = prmt(x1, y1, s.C);
  d = prmt(x1, y1, s.D);
#endif
}

// ---- shift (:243-260) on four lines at once -----------------------------------------
// Compaction (:250-251): every tile moves toward a by the number of empty cells in front of it
// (0..3).  Done as a two-stage logarithmic shifter instead of six dependent conditional swaps:
// stage 1 moves a tile by one place if that number is odd, stage 2 by two places if it is >= 2.
// za, zb, zc = 0xFF in every byte lane where a, b, c is empty.  A destination is always empty
// when a tile arrives (its own tile has left, or it was a hole), so OR inserts the tile.
//   tiles in front of b: za            -> b moves 1 if za
//   in front of c: za + zb             -> c moves 1 if za ^ zb, then 2 if za & zb
//   in front of d: za + zb + zc        -> d moves 1 if za ^ zb ^ zc, then 2 if at least two are set
// (with all three set d has already moved to c in stage 1, and c's rule za & zb carries it on to a;
//  the stage-2 rule for slot d then moves an empty byte, which is harmless).
G2048_DEV void compact(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  const uint32_t za = ~nzmask(a), zb = ~nzmask(b), zc = ~nzmask(c);
  const uint32_t oc = za ^ zb;                     // c moves one place
  const uint32_t od = oc ^ zc;                     // d moves one place
  const uint32_t a1 = a | (b & za);
  const uint32_t b1 = (b & ~za) | (c & oc);
  const uint32_t c1 = (c & ~oc) | (d & od);
  const uint32_t d1 = d & ~od;
  const uint32_t tc = za & zb;                                    // slot c moves on to a
  const uint32_t td = (za & zb) | (za & zc) | (zb & zc);          // slot d moves on to b
  a = a1 | (c1 & tc);
  b = b1 | (d1 & td);
  c = c1 & ~tc;
  d = d1 & ~td;
}

// 2^e (e = byte k of the biased slot word, 0 where the slot is empty) as a float: the byte is
// e+127, moved onto the exponent field (bits 23..30); an empty slot gives +0.0f.
template <int K> G2048_DEV float slot_pow2(uint32_t biased) {
  constexpr uint32_t EXPF = 0x7F800000u;
  uint32_t bits;
  if constexpr (K == 3) bits = shr<1>(biased) & EXPF;
  else if constexpr (K == 0) bits = shl<23>(biased);      // only bit 8 of `biased` survives below the field: the sign
  else bits = shl<23 - 8 * K>(biased) & EXPF;
  float f;
#ifdef G2048_HOST_SIM
  __builtin_memcpy(&f, &bits, 4);
#else
  f = __uint_as_float(bits);
#endif
  return K == 0 ? __builtin_fabsf(f) : f;       // |x| is an operand modifier of FADD: no instruction
}
// A slot word (merged exponents, one per byte lane) with the float exponent bias added where the
// lane holds a merge: e + 127 where filled (<= 145: no carry), 0 elsewhere.
G2048_DEV uint32_t slots_bias(uint32_t slots, uint32_t filled_mask) { return addf(slots, filled_mask & L7); }
// Sum of 2^e over the four lanes of a biased slot word.
G2048_DEV float biased_sum(uint32_t biased) {
  return (slot_pow2<0>(biased) + slot_pow2<1>(biased)) + (slot_pow2<2>(biased) + slot_pow2<3>(biased));
}
// The move score (:254) from the two biased slot words of slide_merge_slots: an exact float (a sum
// of at most 8 powers of two <= 2^18).
G2048_DEV float slots_score(uint32_t biased_a, uint32_t biased_b) {
#if G2048_ABLATE & 8
  return (float)((biased_a ^ biased_b) & 0xFFu);
#endif
  return biased_sum(biased_a) + biased_sum(biased_b);
}

// Slide (a,b,c,d) toward a with merging.  The merged tiles come back as two biased slot words
// (slots_score turns them into the move score): the step kernel carries those two registers from
// the move half of one board into the finishing half, which runs an iteration later.
G2048_DEV void slide_merge_slots(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t& biased_a,
                                 uint32_t& biased_b) {
#if G2048_ABLATE & 128
  biased_a = a & 0x7F; biased_b = b & 0x7F; a ^= d; return;
#endif
  compact(a, b, c, d);
#if G2048_ABLATE & 64
  biased_a = a & 0x7F; biased_b = b & 0x7F; return;
#endif
  // merges on the compacted line: leftmost pair first, each tile merges once (:252-259)
  //   m1: a==b!=0;  m2: b==c!=0 and not m1;  m3: c==d!=0 and not m2   (bit 7 of each byte)
  const uint32_t m1 = ~addf(a ^ b, L7) & addf(b, L7);
  const uint32_t m2 = ~addf(b ^ c, L7) & addf(c, L7) & ~m1;
  const uint32_t m3 = ~addf(c ^ d, L7) & addf(d, L7) & ~m2;
  const uint32_t M1 = spread(m1), M2 = spread(m2), M3 = spread(m3);
  // results written directly in their final places (value * 2 == exponent + 1, :253):
  //   m1&m3: (a+1, c+1, 0, 0)  m1: (a+1, c, d, 0)  m2: (a, b+1, d, 0)  m3: (a, b, c+1, 0)
  const uint32_t cinc = addf(c, M3 & K1);
  const uint32_t binc = addf(b, M2 & K1);
  a = addf(a, M1 & K1);
  b = (cinc & M1) | (binc & ~M1);
  c = (M1 & d & ~M3) | (~M1 & ((M2 & d) | (~M2 & cinc)));
  d &= ~(M1 | M2 | M3);
  // merged exponents: per lane either (m1 and maybe m3) or m2 alone -> two slot words
  const uint32_t sA = (a & M1) | (b & M2);
  const uint32_t sB = cinc & M3;
  biased_a = slots_bias(sA, M1 | M2);
  biased_b = slots_bias(sB, M3);
}
G2048_DEV float slide_merge(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  uint32_t ba, bb;
  slide_merge_slots(a, b, c, d, ba, bb);
  return slots_score(ba, bb);
}

// ---- add_tile (:166-176) under the draw-stream definition ----------------------------
// Places 2 (P=0.9) or 4 on the k-th empty cell (row-major), k = (w * n_empty) >> 32.
// Returns n_empty BEFORE the spawn (0 => nothing placed).
// `enable` = 0xFFFFFFFF to place the tile, 0 to leave the board untouched (illegal move:
// no tile, :91-95).
//
// Build-time flavours (scripts/kernel_variants.py; all bit-identical):
//   G2048_SPAWN_ALU    1: the spawn's adds and constant shifts as IADD3 / SHF (ALU pipe) instead of IMAD /
//                         IMAD.HI: the spawn follows the Philox rounds in the instruction stream and is the
//                         FMA-heavy stretch of the step, while the move before it is the ALU-heavy one
//   G2048_MADHI_PAIR   1: the tile insertion hi32(t * m) + r as ONE wide multiply-add whose 64-bit addend is the
//                         register pair {t, r} — mad.hi's addend {0, r} costs a register zeroing per row; the
//                         low half of t * m is 0 for every t, m used here, so any low addend word is harmless
#ifndef G2048_SPAWN_ALU
#define G2048_SPAWN_ALU 0
#endif
#ifndef G2048_MADHI_PAIR
#define G2048_MADHI_PAIR 0
#endif
#ifdef G2048_HOST_SIM
inline uint32_t sp_add(uint32_t x, uint32_t y) { return x + y; }
template <int S> inline uint32_t sp_shr(uint32_t x) { return x >> S; }
inline uint32_t insert_tile(uint32_t t, uint32_t mult, uint32_t r) { return madhi(t, mult, r); }
#else
__device__ __forceinline__ uint32_t sp_add(uint32_t x, uint32_t y) { return G2048_SPAWN_ALU ? (x + y) : addf(x, y); }
template <int S> __device__ __forceinline__ uint32_t sp_shr(uint32_t x) { return G2048_SPAWN_ALU ? (x >> S) : shr<S>(x); }
__device__ __forceinline__ uint32_t insert_tile(uint32_t t, uint32_t mult, uint32_t r) {
#if G2048_MADHI_PAIR
  uint32_t d;
  asm("{\n\t.reg .b64 c, p;\n\t.reg .b32 lo;\n\t"
      "mov.b64 c, {%1, %3};\n\t"
      "mad.wide.u32 p, %1, %2, c;\n\t"
      "mov.b64 {lo, %0}, p;\n\t}"
      : "=r"(d) : "r"(t), "r"(mult), "r"(r));
  return d;
#else
  return madhi(t, mult, r);
#endif
}
#endif
G2048_DEV uint32_t spawn(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t w,
                         uint32_t enable = 0xFFFFFFFFu) {
#if G2048_ABLATE & 32
  r0 ^= w & enable & 1u; return 2u + (w & 3u);
#endif
  // e_i: bit 7 set where the cell is empty; q_i: 0/1 per byte
  const uint32_t e0 = ~sp_add(r0, L7) & H, e1 = ~sp_add(r1, L7) & H, e2 = ~sp_add(r2, L7) & H, e3 = ~sp_add(r3, L7) & H;
  const uint32_t q0 = sp_shr<7>(e0), q1 = sp_shr<7>(e1), q2 = sp_shr<7>(e2), q3 = sp_shr<7>(e3);
  // inclusive row-major prefix counts of empties, one per byte (<= 16: no carries);
  // byte 3 of each word is the running total, broadcast into the next row's offset
  const uint32_t p0 = q0 * K1;
  const uint32_t p1 = q1 * K1 + prmt(p0, 0u, 0x3333u);
  const uint32_t p2 = q2 * K1 + prmt(p1, 0u, 0x3333u);
  const uint32_t p3 = q3 * K1 + prmt(p2, 0u, 0x3333u);
  const uint32_t n = sp_shr<24>(p3);
  const uint32_t k = __umulhi(w, n);
  const uint32_t f = w * n;
  // target = the cell with prefix == k+1 that is empty:  (p > k) and not (p > k+1)
  const uint32_t gk = L7 - k * K1;          // p + gk has bit 7  <=>  p >= k+1
  const uint32_t gk1 = gk - K1;             // p + gk1 has bit 7 <=>  p >= k+2
  // The target flag t_i has bit 7 of exactly one byte set (e_i is clean).  Shifting it right by 7
  // (tile 2, exponent 1) or 6 (tile 4, exponent 2) gives the tile byte; the cell is empty, so adding
  // equals inserting.  The shift is the high half of t_i * 2^25 / 2^26: one IMAD.HI per row, and a
  // zero multiplier (illegal move: no tile, :91-95) disables the spawn.
#ifndef G2048_SHIFT_INSERT
#define G2048_SHIFT_INSERT 1
#endif
#if G2048_SHIFT_INSERT
  // the flag (bit 7 of the target byte) shifted down to the tile's exponent: >> 7 for a 2, >> 6 for a 4; a shift
  // count of 32 (funnel shift of a zero upper word) drops the flag: no tile after an illegal move (:91-95)
  const uint32_t sh = enable ? ((f < P2_THRESHOLD) ? 7u : 6u) : 32u;                              // :168
  r0 += __funnelshift_rc(sp_add(p0, gk) & ~sp_add(p0, gk1) & e0, 0u, sh);
  r1 += __funnelshift_rc(sp_add(p1, gk) & ~sp_add(p1, gk1) & e1, 0u, sh);
  r2 += __funnelshift_rc(sp_add(p2, gk) & ~sp_add(p2, gk1) & e2, 0u, sh);
  r3 += __funnelshift_rc(sp_add(p3, gk) & ~sp_add(p3, gk1) & e3, 0u, sh);
#elif defined(G2048_SPREAD_INSERT) && G2048_SPREAD_INSERT
  const uint32_t tile = ((f < P2_THRESHOLD) ? K1 : 2u * K1) & enable;                             // :168
  r0 |= spread(sp_add(p0, gk) & ~sp_add(p0, gk1) & e0) & tile;
  r1 |= spread(sp_add(p1, gk) & ~sp_add(p1, gk1) & e1) & tile;
  r2 |= spread(sp_add(p2, gk) & ~sp_add(p2, gk1) & e2) & tile;
  r3 |= spread(sp_add(p3, gk) & ~sp_add(p3, gk1) & e3) & tile;
#else
  const uint32_t mult = ((f < P2_THRESHOLD) ? (1u << 25) : (1u << 26)) & enable;                  // :168
#if G2048_ABLATE & 2
  r0 ^= (sp_add(p0, gk) & ~sp_add(p0, gk1) & e0) ^ (sp_add(p1, gk) & ~sp_add(p1, gk1) & e1) ^ (sp_add(p2, gk) & ~sp_add(p2, gk1) & e2) ^
        (sp_add(p3, gk) & ~sp_add(p3, gk1) & e3) ^ mult;
  return n;
#endif
  r0 = insert_tile(sp_add(p0, gk) & ~sp_add(p0, gk1) & e0, mult, r0);
  r1 = insert_tile(sp_add(p1, gk) & ~sp_add(p1, gk1) & e1, mult, r1);
  r2 = insert_tile(sp_add(p2, gk) & ~sp_add(p2, gk1) & e2, mult, r2);
  r3 = insert_tile(sp_add(p3, gk) & ~sp_add(p3, gk1) & e3, mult, r3);
#endif
  return n;
}

// reset (:102-111): zero board, two spawns from w1, w2.  Specialised: the first spawn sees
// 16 empties (k1 = w1 >> 28), the second 15 and skips cell k1.  A board with one tile is one
// of 32 patterns (16 cells x {2,4}), kept as a uint4 table (shared memory in the kernels):
// the fresh board is the OR of two entries.
struct alignas(16) Board4 { uint32_t x, y, z, w; };
G2048_HD_CONSTEXPR Board4 one_tile_board(uint32_t entry) {        // entry = cell * 2 + (tile == 4)
  const uint32_t cell = entry >> 1, v = ((entry & 1u) + 1u) << ((cell & 3u) * 8u), row = cell >> 2;
  return Board4{row == 0 ? v : 0u, row == 1 ? v : 0u, row == 2 ? v : 0u, row == 3 ? v : 0u};
}
G2048_DEV void fresh_board(const Board4* lut, uint32_t w1, uint32_t w2, uint32_t& r0, uint32_t& r1,
                           uint32_t& r2, uint32_t& r3) {
  const uint32_t k1 = w1 >> 28;
  const uint32_t i1 = 2u * k1 + ((shl<4>(w1) < P2_THRESHOLD) ? 0u : 1u);
  uint32_t k2 = __umulhi(w2, 15u);
  k2 += (k2 >= k1) ? 1u : 0u;
  const uint32_t i2 = 2u * k2 + (((w2 * 15u) < P2_THRESHOLD) ? 0u : 1u);
  const Board4 t1 = lut[i1], t2 = lut[i2];
  r0 = t1.x | t2.x; r1 = t1.y | t2.y; r2 = t1.z | t2.z; r3 = t1.w | t2.w;
}

// The same through a table of whole fresh boards, one per outcome of the two spawns: entry
// k1*64 + k2r*4 + t1*2 + t2, k1 = cell of the first tile, k2r = rank of the second tile's cell among
// the 15 cells left (the table builder skips k1), t = 1 for a 4.  1024 entries (k2r = 15 unused),
// 16 KB of shared memory: half the instructions of fresh_board in the reset path of the step kernel.
G2048_HD_CONSTEXPR Board4 two_tile_board(uint32_t entry) {
  const uint32_t k1 = entry >> 6, k2r = (entry >> 2) & 15u, t1 = (entry >> 1) & 1u, t2 = entry & 1u;
  const uint32_t k2 = (k2r + ((k2r >= k1) ? 1u : 0u)) & 15u;
  const Board4 b1 = one_tile_board(2u * k1 + t1), b2 = one_tile_board(2u * k2 + t2);
  return Board4{b1.x | b2.x, b1.y | b2.y, b1.z | b2.z, b1.w | b2.w};
}
G2048_DEV void fresh_board_pairs(const Board4* pair_lut, uint32_t w1, uint32_t w2, uint32_t& r0, uint32_t& r1,
                                 uint32_t& r2, uint32_t& r3) {
  const uint32_t k1 = w1 >> 28;
  const uint32_t t1 = (shl<4>(w1) < P2_THRESHOLD) ? 0u : 2u;
  const uint32_t k2r = __umulhi(w2, 15u);
  const uint32_t t2 = ((w2 * 15u) < P2_THRESHOLD) ? 0u : 1u;
  const Board4 b = pair_lut[k1 * 64u + k2r * 4u + t1 + t2];
  r0 = b.x; r1 = b.y; r2 = b.z; r3 = b.w;
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
  // (the adds are pinned to the FMA pipe, addm: the ALU pipe is the step kernel's critical one)
  const uint32_t n0 = addm(r0, L7), n1 = addm(r1, L7), n2 = addm(r2, L7), n3 = addm(r3, L7);   // bit7: non-empty
  // vertical pairs (upper u, lower l)
  const uint32_t eqv = (~addm(r0 ^ r1, L7) & n0) | (~addm(r1 ^ r2, L7) & n1) | (~addm(r2 ^ r3, L7) & n2);
  const uint32_t up = (~n0 & n1) | (~n1 & n2) | (~n2 & n3) | eqv;     // hole above a tile
  const uint32_t dn = (n0 & ~n1) | (n1 & ~n2) | (n2 & ~n3) | eqv;     // hole below a tile
  // horizontal pairs: byte j of s_i is cell (i, j+1); only byte lanes 0..2 are pairs
  const uint32_t s0 = shr<8>(r0), s1 = shr<8>(r1), s2 = shr<8>(r2), s3 = shr<8>(r3);
  const uint32_t m0 = addm(s0, L7), m1 = addm(s1, L7), m2 = addm(s2, L7), m3 = addm(s3, L7);
  const uint32_t eqh = (~addm(r0 ^ s0, L7) & n0) | (~addm(r1 ^ s1, L7) & n1) |
                       (~addm(r2 ^ s2, L7) & n2) | (~addm(r3 ^ s3, L7) & n3);
  const uint32_t lf = (~n0 & m0) | (~n1 & m1) | (~n2 & m2) | (~n3 & m3) | eqh;   // hole left of a tile
  const uint32_t rt = (n0 & ~m0) | (n1 & ~m1) | (n2 & ~m2) | (n3 & ~m3) | eqh;   // hole right of a tile
  constexpr uint32_t HP = 0x00808080u;                                           // pair lanes only
  return ((up & H) ? 1u : 0u) | ((rt & HP) ? 2u : 0u) | ((dn & H) ? 4u : 0u) | ((lf & HP) ? 8u : 0u);
}

G2048_DEV uint32_t count_empty(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  return 16u - __popc((r0 + L7) & H) - __popc((r1 + L7) & H) - __popc((r2 + L7) & H) - __popc((r3 + L7) & H);
}

// ---- board symmetries of the reference's training_data.py (:257-279) ------------------------
// hflip: np.flip(x, 2) — every row reversed; swaps actions 1 and 3 (:262-268).
G2048_DEV void board_hflip(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  r0 = prmt(r0, 0u, 0x0123u); r1 = prmt(r1, 0u, 0x0123u); r2 = prmt(r2, 0u, 0x0123u); r3 = prmt(r3, 0u, 0x0123u);
}
G2048_DEV uint32_t action_hflip(uint32_t a) { return (a & 1u) ? (a ^ 2u) : a; }
// rotate(1): np.rot90(x, k=1, axes=(2,1)) on [n,4,4] = a clockwise quarter turn, new[r][c] =
// old[3-c][r]: new row R is column R of the old board read bottom to top; action + 1 mod 4 (:277).
template <int R> G2048_DEV uint32_t rot1_row(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  constexpr uint32_t sel = (uint32_t)R | ((4u + R) << 4);   // byte0 = first[R], byte1 = second[R]
  const uint32_t lo = prmt(r3, r2, sel);                     // old[3][R], old[2][R]
  const uint32_t hi = prmt(r1, r0, sel);                     // old[1][R], old[0][R]
  return prmt(lo, hi, 0x5410u);
}
G2048_DEV void board_rot1(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  const uint32_t n0 = rot1_row<0>(r0, r1, r2, r3), n1 = rot1_row<1>(r0, r1, r2, r3);
  const uint32_t n2 = rot1_row<2>(r0, r1, r2, r3), n3 = rot1_row<3>(r0, r1, r2, r3);
  r0 = n0; r1 = n1; r2 = n2; r3 = n3;
}

// ---- random policies (train.py:119 random.randint(0, 3); random-legal of BASELINE config 4) ---
// The k-th set bit of the 4-bit mask of allowed actions (all four when the mask is empty),
// k = hi32(w * popcount): uniform over the allowed set.
// Branch-free: the index of the k-th set bit of every 4-bit mask comes from a 16-byte table (byte m holds the four
// 2-bit answers for k = 0..3), read with two PRMTs.
G2048_HD_CONSTEXPR uint32_t kth_bit_byte(uint32_t m) {
  uint32_t out = 0, k = 0;
  for (uint32_t b = 0; b < 4; ++b)
    if (m & (1u << b)) out |= b << (2u * k++);
  return out;
}
G2048_HD_CONSTEXPR uint32_t kth_bit_word(uint32_t m0) {
  return kth_bit_byte(m0) | (kth_bit_byte(m0 + 1) << 8) | (kth_bit_byte(m0 + 2) << 16) | (kth_bit_byte(m0 + 3) << 24);
}
G2048_DEV uint32_t pick_action(uint32_t mask, uint32_t w) {
  uint32_t m = mask & 15u;
  if (m == 0u) m = 15u;
  const uint32_t k = __umulhi(w, __popc(m));
  constexpr uint32_t T0 = kth_bit_word(0), T1 = kth_bit_word(4), T2 = kth_bit_word(8), T3 = kth_bit_word(12);
  const uint32_t lo = prmt(T0, T1, m & 7u), hi = prmt(T2, T3, m & 7u);      // byte 0 = table[m & 7] / table[8 + (m & 7)]
  const uint32_t byte = (m & 8u) ? hi : lo;
  return (byte >> (2u * k)) & 3u;
}

// ---- one whole step (:76-100) on a board held in registers --------------------------
struct StepOut {
  float score;        // merge score of the move (0 when illegal), exact
  uint32_t highest;   // exponent of the highest tile after the spawn (valid if requested)
  bool legal;         // False <=> IllegalMove (:91-95)
  bool done;          // terminated
  uint32_t t0, t1, t2, t3;   // post-spawn board (the terminal board when done)
};

// The step comes in two halves so that the step kernel can run them on DIFFERENT boards in the
// same loop iteration (software pipelining: the move half is ALU-pipe work — PRMT/LOP3 —, the
// finishing half FMA-pipe work — Philox, prefix counts, float score; interleaved they keep both
// pipes busy, one after the other they take turns idling).
//
// Move half: a board already rotated into the move frame (a,b,c,d) (see orient; `so` is the
// inverse rotation) through move() (:194-241).  What the finishing half needs is 7 registers.
struct Moved {
  uint32_t r0, r1, r2, r3;     // the moved board, back in row-major order (unchanged when illegal)
  uint32_t biased_a, biased_b; // merged tiles (slots_score)
  uint32_t legal;              // 0xFFFFFFFF, or 0 <=> IllegalMove (:238-239): nothing changed
};
G2048_DEV Moved move_oriented(uint32_t a, uint32_t b, uint32_t c, uint32_t d, const Sel4 so) {
  Moved m;
  const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
  slide_merge_slots(a, b, c, d, m.biased_a, m.biased_b);                                  // :85, :194-241
  // An illegal move leaves (a,b,c,d) as they were and scores 0, so the same data path serves
  // both cases; only the spawn is gated.
  m.legal = ((((a ^ a0) | (b ^ b0)) | ((c ^ c0) | (d ^ d0))) != 0u) ? 0xFFFFFFFFu : 0u;
  orient(so, a, b, c, d, m.r0, m.r1, m.r2, m.r3);
  return m;
}

// Finishing half: spawn, score, isend, auto-reset.  On return r0..r3 hold the board handed back
// to the agent: the post-spawn board, or a fresh reset() board when the episode ended and
// auto_reset is set (SB3 DummyVecEnv).
template <bool PAIR_LUT = false>
G2048_DEV StepOut finish_step(const Board4* lut, const Moved& m, const Words& w, uint32_t max_tile_exp,
                              bool want_highest, bool auto_reset, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                              uint32_t& r3) {
  StepOut o;
  o.legal = m.legal != 0u;
  o.score = slots_score(m.biased_a, m.biased_b);
  r0 = m.r0; r1 = m.r1; r2 = m.r2; r3 = m.r3;
  const uint32_t n_empty = spawn(r0, r1, r2, r3, w.w0, m.legal);                         // :88
  o.highest = 0;
  if (want_highest || max_tile_exp != 0u) o.highest = highest_exp(r0, r1, r2, r3);      // :97
  // isend (:262-280) for a legal move; an illegal move terminates (:94)
  bool end = (n_empty == 1u) && full_board_is_dead(r0, r1, r2, r3);
  if (max_tile_exp != 0u) end = end || (o.highest == max_tile_exp);                      // :267
  o.done = end || !o.legal;
  o.t0 = r0; o.t1 = r1; o.t2 = r2; o.t3 = r3;
  if (auto_reset && o.done) {                                                            // :102-111
    if (PAIR_LUT) fresh_board_pairs(lut, w.w1, w.w2, r0, r1, r2, r3);
    else fresh_board(lut, w.w1, w.w2, r0, r1, r2, r3);
  }
  return o;
}

G2048_DEV StepOut step_oriented(const Board4* lut, uint32_t a, uint32_t b, uint32_t c, uint32_t d, const Sel4 so,
                                const Words& w, uint32_t max_tile_exp, bool want_highest, bool auto_reset,
                                uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  const Moved m = move_oriented(a, b, c, d, so);
  return finish_step(lut, m, w, max_tile_exp, want_highest, auto_reset, r0, r1, r2, r3);
}

G2048_DEV StepOut step_board(const Board4* lut, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                             uint32_t action, const Words& w, uint32_t max_tile_exp, bool want_highest,
                             bool auto_reset) {
  uint32_t a, b, c, d;
  orient(kOrientIn[action], r0, r1, r2, r3, a, b, c, d);
  return step_oriented(lut, a, b, c, d, kOrientOut[action], w, max_tile_exp, want_highest, auto_reset, r0, r1, r2,
                       r3);
}

}  // namespace g2048
