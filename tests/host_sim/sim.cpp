// Host-side simulation of the CUDA step logic — TEST ONLY.
//
// Compiles gym-2048_b200/csrc/g2048_device.cuh with g++ (G2048_HOST_SIM: prmt/umulhi/popc
// emulated) so the exact byte-SIMD code the kernel runs can be checked against the oracle
// and the golden vectors on the GPU-less build box.  Never loaded by the product.
#define G2048_HOST_SIM 1
#include <cstring>
#include "../../gym-2048_b200/csrc/g2048_device.cuh"
#include "../../include/g2048.h"

using namespace g2048;

static Board4 g_lut[32];
static const Board4* lut() {
  static bool init = false;
  if (!init) { for (uint32_t e = 0; e < 32; ++e) g_lut[e] = one_tile_board(e); init = true; }
  return g_lut;
}

static inline void load(const uint8_t* b, uint32_t r[4]) { std::memcpy(r, b, 16); }
static inline void store(uint8_t* b, const uint32_t r[4]) { std::memcpy(b, r, 16); }

extern "C" int sim_step(const G2048StepArgs* p) {
  for (uint64_t i = 0; i < p->n; ++i) {
    uint32_t r[4];
    load(p->boards + 16 * i, r);
    Words w;
    if (p->forced_draws) {
      const uint32_t* f = p->forced_draws + 4 * i;
      w = Words{f[0], f[1], f[2], f[3]};
    } else {
      // the kernel's hot-loop form: host-made round keys with idx_lo folded into round 0
      const uint64_t env = p->env_id_base + i, idx = p->step_counter ? *p->step_counter : p->step_index;
      StreamKeys ks;
      make_stream_keys(stream_key(p->seed, idx, env, TAG_STEP), (uint32_t)idx, ks);
      w = words_from_pair(philox2x32_10_keys((uint32_t)env, ks));
    }
    const bool auto_reset = (p->flags & G2048_FLAG_AUTO_RESET) != 0;
    StepOut o = step_board(lut(), r[0], r[1], r[2], r[3], p->actions[i] & 3u, w, p->max_tile_exp,
                           p->highest_exp != nullptr, auto_reset);
    store(p->boards + 16 * i, r);
    p->rewards[i] = o.legal ? o.score : p->illegal_move_reward;
    p->dones[i] = o.done;
    if (p->illegal) p->illegal[i] = !o.legal;
    if (p->highest_exp) p->highest_exp[i] = (uint8_t)o.highest;
    uint32_t es = p->ep_score ? p->ep_score[i] + (uint32_t)o.score : 0, el = p->ep_len ? p->ep_len[i] + 1 : 0;
    float er = p->ep_return ? p->ep_return[i] + p->rewards[i] : 0.f;
    if (o.done) {
      const uint32_t t[4] = {o.t0, o.t1, o.t2, o.t3};
      if (p->terminal_boards) store(p->terminal_boards + 16 * i, t);
      if (p->final_score) p->final_score[i] = es;
      if (p->final_len) p->final_len[i] = el;
      if (p->final_return) p->final_return[i] = er;
      if (auto_reset) { es = el = 0; er = 0.f; }
    }
    if (p->ep_score) p->ep_score[i] = es;
    if (p->ep_len) p->ep_len[i] = el;
    if (p->ep_return) p->ep_return[i] = er;
    if (p->legal_mask) p->legal_mask[i] = (uint8_t)legal_mask(r[0], r[1], r[2], r[3]);
  }
  if (p->step_counter) *p->step_counter += 1;
  return 0;
}

extern "C" int sim_reset(uint8_t* boards, const uint8_t* mask, uint64_t n, uint64_t base, uint64_t seed,
                         uint64_t reset_index) {
  for (uint64_t i = 0; i < n; ++i) {
    if (mask && !mask[i]) continue;
    Words w = draw_words(seed, base + i, reset_index, TAG_RESET);
    uint32_t r[4];
    fresh_board(lut(), w.w1, w.w2, r[0], r[1], r[2], r[3]);
    store(boards + 16 * i, r);
  }
  return 0;
}

extern "C" int sim_add_tile(uint8_t* boards, uint64_t n, uint64_t base, uint64_t seed, uint64_t step_index) {
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t r[4];
    load(boards + 16 * i, r);
    Words w = draw_words(seed, base + i, step_index, TAG_STEP);
    spawn(r[0], r[1], r[2], r[3], w.w0);
    store(boards + 16 * i, r);
  }
  return 0;
}

extern "C" int sim_move(const uint8_t* in, uint8_t* out, const uint8_t* dirs, uint32_t* scores,
                        uint8_t* changed, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t r[4], a, b, c, d;
    load(in + 16 * i, r);
    const uint32_t act = dirs[i] & 3u;
    orient(kOrientIn[act], r[0], r[1], r[2], r[3], a, b, c, d);
    const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
    const uint32_t s = (uint32_t)slide_merge(a, b, c, d);
    orient(kOrientOut[act], a, b, c, d, r[0], r[1], r[2], r[3]);
    if (out) store(out + 16 * i, r);
    if (scores) scores[i] = s;
    if (changed) changed[i] = ((a ^ a0) | (b ^ b0) | (c ^ c0) | (d ^ d0)) != 0;
  }
  return 0;
}

extern "C" int sim_status(const uint8_t* boards, uint8_t* lm, uint8_t* hi, uint8_t* ne, uint8_t* end,
                          uint32_t max_tile_exp, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t r[4];
    load(boards + 16 * i, r);
    const uint32_t empties = count_empty(r[0], r[1], r[2], r[3]);
    const uint32_t h = highest_exp(r[0], r[1], r[2], r[3]);
    if (lm) lm[i] = (uint8_t)legal_mask(r[0], r[1], r[2], r[3]);
    if (hi) hi[i] = (uint8_t)h;
    if (ne) ne[i] = (uint8_t)empties;
    if (end) end[i] = (max_tile_exp != 0 && h == max_tile_exp) ||
                      (empties == 0 && full_board_is_dead(r[0], r[1], r[2], r[3]));
  }
  return 0;
}

extern "C" int sim_philox(const uint32_t* ctr, uint32_t k0, uint32_t k1, uint32_t* out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    Words w = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], k0, k1);
    out[4 * i] = w.w0; out[4 * i + 1] = w.w1; out[4 * i + 2] = w.w2; out[4 * i + 3] = w.w3;
  }
  return 0;
}

extern "C" int sim_philox2x32(const uint32_t* ctr, uint32_t key, uint32_t* out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    const Pair x = philox2x32_10(ctr[2 * i], ctr[2 * i + 1], key);
    out[2 * i] = x.x0; out[2 * i + 1] = x.x1;
  }
  return 0;
}

// The three forms the kernels use must agree: draw_words (reset/add_tile/policy kernels, one
// board at a time), DrawStream (the same with the key cached across boards) and the step
// kernel's precomputed round keys.  form: 0, 1, 2.
extern "C" int sim_draw_words(uint32_t* out, uint64_t n, uint64_t base, uint64_t seed, uint64_t idx, uint32_t tag,
                              int form) {
  DrawStream ds(seed, idx, tag);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t env = base + i;
    Words w;
    if (form == 0) w = draw_words(seed, env, idx, tag);
    else if (form == 1) w = ds.words(env);
    else {
      StreamKeys ks;
      make_stream_keys(stream_key(seed, idx, env, tag), (uint32_t)idx, ks);
      w = words_from_pair(philox2x32_10_keys((uint32_t)env, ks));
    }
    out[4 * i] = w.w0; out[4 * i + 1] = w.w1; out[4 * i + 2] = w.w2; out[4 * i + 3] = w.w3;
  }
  return 0;
}

extern "C" int sim_symmetry(const uint8_t* in, uint8_t* out, const uint8_t* act_in, uint8_t* act_out, uint64_t n,
                            int hflip, int k) {
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t r[4];
    load(in + 16 * i, r);
    if (hflip) board_hflip(r[0], r[1], r[2], r[3]);
    for (int q = 0; q < k; ++q) board_rot1(r[0], r[1], r[2], r[3]);
    store(out + 16 * i, r);
    if (act_in && act_out) {
      uint32_t a = act_in[i] & 3u;
      if (hflip) a = action_hflip(a);
      act_out[i] = (uint8_t)((a + (uint32_t)k) & 3u);
    }
  }
  return 0;
}

extern "C" int sim_sample_actions(const uint8_t* mask, uint8_t* actions, uint64_t n, uint64_t base, uint64_t seed,
                                  uint64_t step_index) {
  for (uint64_t i = 0; i < n; ++i) {
    Words w = draw_words(seed, base + i, step_index, TAG_POLICY);
    actions[i] = (uint8_t)pick_action(mask ? mask[i] : 15u, w.w0);
  }
  return 0;
}

// reset through the 1024-entry fresh-board table (what the step kernels read) and through the 32-entry
// one-tile table (reset kernel, oracle-checked): boards_pairs / boards_lut, 16 bytes each
extern "C" int sim_fresh_boards(const uint32_t* w1, const uint32_t* w2, uint8_t* boards_pairs, uint8_t* boards_lut,
                                uint64_t n) {
  static Board4 table[1024];
  static bool init = false;
  if (!init) { for (uint32_t e = 0; e < 1024; ++e) table[e] = two_tile_board(e); init = true; }
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t r[4];
    fresh_board_pairs(table, w1[i], w2[i], r[0], r[1], r[2], r[3]);
    store(boards_pairs + 16 * i, r);
    fresh_board(lut(), w1[i], w2[i], r[0], r[1], r[2], r[3]);
    store(boards_lut + 16 * i, r);
  }
  return 0;
}

