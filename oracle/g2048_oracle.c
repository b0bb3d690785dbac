/*
 * g2048_oracle.c — CPU restatement of rgal/gym-2048's Game2048Env hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker or the timed CPU baseline.  The product (libg2048.so,
 * gym-2048_b200/) never links, imports or falls back to it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * (i) the known answers of the reference's own env/envs/test_game2048_env.py,
 * (ii) the 848 transitions of the reference's data/test_data.csv and (iii)
 * outputs of the UNMODIFIED reference env run in the build container with the
 * draw-injection shim (tests/golden/make_golden.py), all committed under
 * tests/golden/.
 *
 * The code follows the reference's scalar algorithm (one Python-style pass per
 * line, four trial moves for isend), NOT the byte-SIMD formulation of the CUDA
 * kernel, so that the two implementations are independent.  Boards are 16 bytes
 * of tile EXPONENTS (value 2^e, 0 = empty) — see include/g2048.h.  Citations
 * are file:line in /root/reference/env/envs/game2048_env.py.
 */
#include <stdint.h>
#include <string.h>
#include "../include/g2048.h"

#include <pthread.h>

/* ---- Philox4x32-10 (Salmon et al., SC'11; Random123 constants) ------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- Philox2x32-10 (same paper, same constants table) --------------------- */
static void philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
  uint32_t c0 = ctr[0], c1 = ctr[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p = (uint64_t)0xD256D193u * c0;
    uint32_t n0 = (uint32_t)(p >> 32) ^ key ^ c1;
    c1 = (uint32_t)p;
    c0 = n0;
    key += 0x9E3779B9u;
  }
  out[0] = c0; out[1] = c1;
}

/* Draw stream version 2 (include/g2048.h): a launch-uniform 32-bit key from
 * Philox4x32-10, then one Philox2x32-10 block per (env, index). */
static void draw_words(uint64_t seed, uint64_t env_id, uint64_t idx, uint32_t tag,
                       uint32_t w[4]) {
  uint32_t kctr[4] = {(uint32_t)(idx >> 32), (uint32_t)(env_id >> 32), tag, 0u};
  uint32_t kkey[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t kout[4], x[2];
  philox4x32_10(kctr, kkey, kout);
  uint32_t ctr[2] = {(uint32_t)env_id, (uint32_t)idx};
  philox2x32_10(ctr, kout[0], x);
  w[0] = x[0]; w[1] = x[1]; w[2] = x[1] << 16; w[3] = 0u;
}

/* ---- shift (:243-260): single pass over one line toward index 0 ----------- */
static uint32_t shift_line(const uint8_t row[4], uint8_t combined[4]) {
  uint32_t move_score = 0;
  int output_index = 0, can_merge = 0;
  combined[0] = combined[1] = combined[2] = combined[3] = 0;
  for (int i = 0; i < 4; ++i) {
    uint8_t val = row[i];
    if (val == 0) continue;                                       /* :250-251 */
    if (can_merge && combined[output_index - 1] == val) {         /* :252 */
      combined[output_index - 1] = (uint8_t)(val + 1);            /* :253 value *= 2 */
      move_score += 1u << (val + 1);                              /* :254 */
      can_merge = 0;                                              /* :255 */
    } else {
      combined[output_index++] = val;                             /* :257-258 */
      can_merge = 1;                                              /* :259 */
    }
  }
  return move_score;
}

/* ---- move (:194-241): returns changed (0 = IllegalMove), *score = move score */
static int move_board(uint8_t b[16], int direction, int trial, uint32_t* score) {
  int changed = 0;
  uint32_t move_score = 0;
  int dir_div_two = direction / 2;                                /* :210 */
  int dir_mod_two = direction % 2;                                /* :211 */
  int shift_direction = dir_mod_two ^ dir_div_two;                /* :212 */
  for (int line = 0; line < 4; ++line) {
    uint8_t old[4], neu[4];
    for (int i = 0; i < 4; ++i) {
      int j = shift_direction ? 3 - i : i;                        /* :218-219, :230-231 */
      old[i] = dir_mod_two == 0 ? b[4 * j + line]                 /* :217 Matrix[:, y] */
                                : b[4 * line + j];                /* :229 Matrix[x, :] */
    }
    move_score += shift_line(old, neu);                           /* :220-221, :232-233 */
    if (memcmp(old, neu, 4) != 0) {                               /* :222, :234 */
      changed = 1;
      if (!trial) {
        for (int i = 0; i < 4; ++i) {
          int j = shift_direction ? 3 - i : i;                    /* :225, :237 */
          if (dir_mod_two == 0) b[4 * j + line] = neu[i];
          else                  b[4 * line + j] = neu[i];
        }
      }
    }
  }
  *score = move_score;
  return changed;                                                 /* :238-239 */
}

/* ---- add_tile (:166-176) under the injected-draw definition (g2048.h) ----- */
static void spawn(uint8_t b[16], uint32_t w) {
  uint32_t n = 0;
  for (int i = 0; i < 16; ++i) n += (b[i] == 0);
  if (n == 0) return;                       /* :176 assert — unreachable from step */
  uint64_t p = (uint64_t)w * n;
  uint32_t k = (uint32_t)(p >> 32);
  uint32_t f = (uint32_t)p;
  uint8_t e = (f < G2048_P2_THRESHOLD) ? 1 : 2;                   /* :168 */
  for (int i = 0; i < 16; ++i) {                                  /* :171-175 */
    if (b[i] == 0) {
      if (k == 0) { b[i] = e; return; }
      --k;
    }
  }
}

static uint8_t highest_exp(const uint8_t b[16]) {                 /* :190-192 */
  uint8_t m = 0;
  for (int i = 0; i < 16; ++i) if (b[i] > m) m = b[i];
  return m;
}

static uint8_t legal_mask_of(const uint8_t b[16]) {               /* :224,236-239 */
  uint8_t mask = 0, tmp[16];
  uint32_t s;
  for (int d = 0; d < 4; ++d) {
    memcpy(tmp, b, 16);
    if (move_board(tmp, d, 1, &s)) mask |= (uint8_t)(1u << d);
  }
  return mask;
}

/* ---- isend (:262-280) ------------------------------------------------------ */
static int isend(const uint8_t b[16], uint32_t max_tile_exp) {
  if (max_tile_exp != 0 && highest_exp(b) == max_tile_exp) return 1;   /* :267 */
  for (int i = 0; i < 16; ++i) if (b[i] == 0) return 0;               /* :270-271 */
  uint8_t tmp[16];
  uint32_t s;
  for (int d = 0; d < 4; ++d) {                                        /* :273-279 */
    memcpy(tmp, b, 16);
    if (move_board(tmp, d, 1, &s)) return 0;
  }
  return 1;                                                            /* :280 */
}

static void reset_board(uint8_t b[16], const uint32_t w[4]) {     /* :102-111 */
  memset(b, 0, 16);                                               /* :104 */
  spawn(b, w[1]);                                                 /* :108 */
  spawn(b, w[2]);                                                 /* :109 */
}

/* ---- step (:76-100) + SB3 DummyVecEnv same-step auto-reset ----------------- */
static void step_one(const G2048StepArgs* a, uint64_t i) {
  uint8_t* b = a->boards + 16 * i;
  if (a->boards_out && a->boards_out != a->boards) {              /* out-of-place step */
    memcpy(a->boards_out + 16 * i, b, 16);
    b = a->boards_out + 16 * i;
  }
  int action = a->actions[i] & 3;
  uint32_t w[4];
  if (a->forced_draws) memcpy(w, a->forced_draws + 4 * i, sizeof w);
  else draw_words(a->seed, a->env_id_base + i, a->step_counter ? *a->step_counter : a->step_index, 0, w);

  uint32_t score = 0;
  float reward;
  int terminated, illegal;
  if (move_board(b, action, 0, &score)) {                         /* :85 */
    illegal = 0;
    spawn(b, w[0]);                                               /* :88 */
    terminated = isend(b, a->max_tile_exp);                       /* :89 */
    reward = (float)score;                                        /* :90 */
  } else {                                                        /* :91-95 */
    illegal = 1;
    terminated = 1;
    reward = a->illegal_move_reward;
    score = 0;
  }
  uint32_t es = 0, el = 0;
  float er = 0.f;
  if (a->ep_score) es = a->ep_score[i] + score;                   /* :86 */
  if (a->ep_len) el = a->ep_len[i] + 1;
  if (a->ep_return) er = a->ep_return[i] + reward;                /* SB3 Monitor.step: rewards.append(reward); sum at done (ppo_train.py:123) */
  a->rewards[i] = reward;
  a->dones[i] = (uint8_t)terminated;
  if (a->illegal) a->illegal[i] = (uint8_t)illegal;
  if (a->highest_exp) a->highest_exp[i] = highest_exp(b);         /* :97 */
  if (terminated) {
    if (a->terminal_boards) memcpy(a->terminal_boards + 16 * i, b, 16);
    if (a->final_score) a->final_score[i] = es;
    if (a->final_len) a->final_len[i] = el;
    if (a->final_return) a->final_return[i] = er;
    if (a->flags & G2048_FLAG_AUTO_RESET) {
      reset_board(b, w);
      es = 0; el = 0; er = 0.f;
    }
  }
  if (a->ep_score) a->ep_score[i] = es;
  if (a->ep_len) a->ep_len[i] = el;
  if (a->ep_return) a->ep_return[i] = er;
  if (a->legal_mask) a->legal_mask[i] = legal_mask_of(b);
}

int g2048_oracle_step(const G2048StepArgs* a) {
  if (!a || !a->boards || !a->actions || !a->rewards || !a->dones) return G2048_ERR_INVALID;
  for (uint64_t i = 0; i < a->n; ++i) step_one(a, i);
  if (a->step_counter) *a->step_counter += 1;
  return G2048_OK;
}

/* Same, split into contiguous slices over `threads` pthreads (bench CPU baseline). */
typedef struct { const G2048StepArgs* a; uint64_t lo, hi; } StepSlice;
static void* step_slice(void* p) {
  StepSlice* s = (StepSlice*)p;
  for (uint64_t i = s->lo; i < s->hi; ++i) step_one(s->a, i);
  return 0;
}
int g2048_oracle_step_mt(const G2048StepArgs* a, int threads) {
  if (!a || !a->boards || !a->actions || !a->rewards || !a->dones) return G2048_ERR_INVALID;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t tid[256];
  StepSlice sl[256];
  uint64_t per = (a->n + (uint64_t)threads - 1) / (uint64_t)threads;
  int started = 0;
  for (int t = 0; t < threads; ++t) {
    sl[t].a = a;
    sl[t].lo = per * (uint64_t)t < a->n ? per * (uint64_t)t : a->n;
    sl[t].hi = sl[t].lo + per < a->n ? sl[t].lo + per : a->n;
    if (pthread_create(&tid[t], 0, step_slice, &sl[t]) != 0) break;
    ++started;
  }
  for (int t = 0; t < started; ++t) pthread_join(tid[t], 0);
  for (int t = started; t < threads; ++t) step_slice(&sl[t]);
  if (a->step_counter) *a->step_counter += 1;
  return G2048_OK;
}

int g2048_oracle_reset(uint8_t* boards, const uint8_t* reset_mask, uint64_t n,
                       uint64_t env_id_base, uint64_t seed, uint64_t reset_index) {
  if (!boards) return G2048_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    if (reset_mask && !reset_mask[i]) continue;
    uint32_t w[4];
    draw_words(seed, env_id_base + i, reset_index, 1, w);
    reset_board(boards + 16 * i, w);
  }
  return G2048_OK;
}

int g2048_oracle_add_tile(uint8_t* boards, uint64_t n, uint64_t env_id_base, uint64_t seed,
                          uint64_t step_index) {
  if (!boards) return G2048_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t w[4];
    draw_words(seed, env_id_base + i, step_index, 0, w);
    spawn(boards + 16 * i, w[0]);
  }
  return G2048_OK;
}

int g2048_oracle_move(const uint8_t* boards_in, uint8_t* boards_out,
                      const uint8_t* directions, uint32_t* scores, uint8_t* changed,
                      uint64_t n) {
  if (!boards_in || !directions) return G2048_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    uint8_t tmp[16];
    uint32_t s;
    memcpy(tmp, boards_in + 16 * i, 16);
    int ch = move_board(tmp, directions[i] & 3, 0, &s);
    if (boards_out) memcpy(boards_out + 16 * i, tmp, 16);
    if (scores) scores[i] = s;
    if (changed) changed[i] = (uint8_t)ch;
  }
  return G2048_OK;
}

int g2048_oracle_status(const uint8_t* boards, uint8_t* legal_mask, uint8_t* hi,
                        uint8_t* n_empty, uint8_t* is_end, uint32_t max_tile_exp,
                        uint64_t n) {
  if (!boards) return G2048_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    const uint8_t* b = boards + 16 * i;
    if (legal_mask) legal_mask[i] = legal_mask_of(b);
    if (hi) hi[i] = highest_exp(b);
    if (n_empty) { uint8_t c = 0; for (int j = 0; j < 16; ++j) c += (b[j] == 0); n_empty[i] = c; }
    if (is_end) is_end[i] = (uint8_t)isend(b, max_tile_exp);
  }
  return G2048_OK;
}

/* stack (:17-32): obs[n][16][4][4] as uint8; ch 0 = empty, ch k = (cell == 2^k). */
int g2048_oracle_encode_obs_u8(const uint8_t* boards, uint8_t* obs, uint64_t n) {
  if (!boards || !obs) return G2048_ERR_INVALID;
  memset(obs, 0, n * 256);
  for (uint64_t i = 0; i < n; ++i)
    for (int cell = 0; cell < 16; ++cell) {
      uint8_t e = boards[16 * i + cell];
      if (e <= 15) obs[256 * i + 16 * e + cell] = 1;   /* e==0 -> channel 0 (:25) */
    }
  return G2048_OK;
}

int g2048_oracle_philox(const uint32_t* ctr, uint32_t key0, uint32_t key1, uint32_t* out,
                        uint64_t n) {
  uint32_t key[2] = {key0, key1};
  for (uint64_t i = 0; i < n; ++i) philox4x32_10(ctr + 4 * i, key, out + 4 * i);
  return G2048_OK;
}

int g2048_oracle_philox2x32(const uint32_t* ctr, uint32_t key, uint32_t* out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) philox2x32_10(ctr + 2 * i, key, out + 2 * i);
  return G2048_OK;
}

int g2048_oracle_draw_words(uint32_t* words, uint64_t n, uint64_t env_id_base, uint64_t seed,
                            uint64_t index, uint32_t tag) {
  for (uint64_t i = 0; i < n; ++i) draw_words(seed, env_id_base + i, index, tag, words + 4 * i);
  return G2048_OK;
}

/* shift() on one line of exponents, for the exhaustive golden table. */
uint32_t g2048_oracle_shift(const uint8_t row[4], uint8_t out[4]) {
  return shift_line(row, out);
}


/* ---- the random policies that drive the step (train.py:119 random.randint(0, 3); ------------
 *      BASELINE config 4's random-legal policy): k-th allowed action, k from word 0 of the
 *      policy-tag stream ---- */
int g2048_oracle_sample_actions(const uint8_t* legal_mask, uint8_t* actions, uint64_t n,
                                uint64_t env_id_base, uint64_t seed, uint64_t step_index) {
  if (!actions) return G2048_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t w[4];
    draw_words(seed, env_id_base + i, step_index, G2048_TAG_POLICY, w);
    int allowed[4], cnt = 0;
    for (int d = 0; d < 4; ++d)
      if (!legal_mask || (legal_mask[i] & 15) == 0 || ((legal_mask[i] >> d) & 1)) allowed[cnt++] = d;
    uint32_t k = (uint32_t)(((uint64_t)w[0] * (uint64_t)cnt) >> 32);
    actions[i] = (uint8_t)allowed[k];
  }
  return G2048_OK;
}

/* ---- training_data.py symmetries, following the numpy calls cell by cell ------------------- */
/* hflip (:257-272): np.flip(x, 2): new[r][c] = old[r][3-c]; actions 1 <-> 3 */
static void hflip_board(const uint8_t* in, uint8_t* out) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[4 * r + c] = in[4 * r + (3 - c)];
}
/* rotate(1) (:274-279): np.rot90(x, k=1, axes=(2,1)).  rot90 with axes (a,b) is
 * flip(transpose) for k=1: result = transpose(flip(m, axis=b)); for one board with
 * (a,b) = (cols, rows): new[r][c] = old[3-c][r]. */
static void rot1_board(const uint8_t* in, uint8_t* out) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[4 * r + c] = in[4 * (3 - c) + r];
}
static void sym_board(const uint8_t* in, uint8_t* out, int hflip, int k) {
  uint8_t t[16], u[16];
  memcpy(t, in, 16);
  if (hflip) { hflip_board(t, u); memcpy(t, u, 16); }
  for (int i = 0; i < k; ++i) { rot1_board(t, u); memcpy(t, u, 16); }
  memcpy(out, t, 16);
}
static uint8_t sym_action(uint8_t a, int hflip, int k) {
  if (hflip) { if (a == 1) a = 3; else if (a == 3) a = 1; }           /* :262-268 */
  return (uint8_t)((a + k) % 4);                                      /* :277 */
}
int g2048_oracle_symmetry(const uint8_t* boards_in, uint8_t* boards_out, const uint8_t* actions_in,
                          uint8_t* actions_out, uint64_t n, int hflip, int k) {
  for (uint64_t i = 0; i < n; ++i) {
    sym_board(boards_in + 16 * i, boards_out + 16 * i, hflip, k);
    if (actions_in && actions_out) actions_out[i] = sym_action(actions_in[i], hflip, k);
  }
  return G2048_OK;
}
/* augment (:281-299): [X, H] then rotations 1..3 of those 2n rows appended */
int g2048_oracle_augment(const uint8_t* boards, const uint8_t* next_boards, const uint8_t* actions,
                         const float* rewards, const uint8_t* dones, uint64_t n, uint8_t* boards_out,
                         uint8_t* next_out, uint8_t* actions_out, float* rewards_out, uint8_t* dones_out) {
  for (int k = 0; k < 4; ++k)
    for (int h = 0; h < 2; ++h)
      for (uint64_t i = 0; i < n; ++i) {
        uint64_t o = (uint64_t)(2 * k + h) * n + i;
        sym_board(boards + 16 * i, boards_out + 16 * o, h, k);
        sym_board(next_boards + 16 * i, next_out + 16 * o, h, k);
        actions_out[o] = sym_action(actions[i], h, k);
        rewards_out[o] = rewards[i];
        dones_out[o] = dones[i];
      }
  return G2048_OK;
}
/* get_discounted_return (:104-124): one reverse pass, `previous` cleared at done rows */
int g2048_oracle_discounted_return(const float* rewards, const uint8_t* dones, double* returns,
                                   uint64_t n, double gamma) {
  int have_prev = 0;
  volatile double previous = 0.0;          /* volatile: keep the two roundings (no fma) */
  for (uint64_t i = n; i-- > 0;) {
    double smoothed = (double)rewards[i];
    if (dones[i]) have_prev = 0;
    if (have_prev && previous != 0.0) {
      volatile double prod = gamma * previous;
      smoothed += prod;
    }
    returns[i] = smoothed;
    previous = smoothed;
    have_prev = 1;
  }
  return G2048_OK;
}
