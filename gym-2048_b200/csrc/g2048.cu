// g2048.cu — CUDA kernels (sm_100a) and the C ABI of libg2048.so (include/g2048.h).
//
// The hot path is g2048_step_kernel: one board per lane, the 16-byte board held in four
// registers, one coalesced 128-bit load and store per board, every game rule evaluated
// branch-free with byte-SIMD integer ops (g2048_device.cuh).  Pure integer/indexing work:
// HBM- and issue-bound, no tensor cores.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#endif
#include <new>

#include "../../include/g2048.h"
#include "g2048_device.cuh"
#include "g2048_internal.h"

namespace g2048 {

// ------------------------------------------------------------------------------------
// error reporting
// ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(G2048_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

// ------------------------------------------------------------------------------------
// launch geometry: grid-stride kernels in one persistent wave over the SMs — kThreads-wide CTAs,
// kCtasPerSm of them per SM (four for the low-register streaming kernels, grid_for_streaming); the
// step kernels have their own shape (kStepThreads x kStepCtasPerSm, shape_for).
// ------------------------------------------------------------------------------------
// Tunables (overridable with -D for scripts/kernel_variants.py experiments)
#ifndef G2048_CTAS_PER_SM
#define G2048_CTAS_PER_SM 2
#endif
#ifndef G2048_PDL            // programmatic dependent launch: overlap this launch's ramp with
#define G2048_PDL 1          // the previous kernel's tail (griddepcontrol.wait guards all loads)
#endif
#ifndef G2048_PREFETCH       // load the next iteration's board/action before computing this one
#define G2048_PREFETCH 1
#endif
#ifndef G2048_PRE_PREFETCH   // L2-prefetch every thread's first board/action BEFORE griddepcontrol.wait (see the kernel)
#define G2048_PRE_PREFETCH 1
#endif
#ifndef G2048_STAGGER        // ns between the start of consecutive warps of an SM sub-partition (0 = off)
#define G2048_STAGGER 0
#endif
#ifndef G2048_PERSISTENT     // 1: grid-stride loop over a grid sized to the SM count; 0: one board per thread
#define G2048_PERSISTENT 1
#endif
#ifndef G2048_TMA            // 1: boards and actions reach the SM through a shared-memory ring filled by bulk async
#define G2048_TMA 0          //    copies (TMA) under mbarriers; 0: per-thread LDG.128 one iteration ahead.
#endif                       //    Measured (profiles/r01_variants_v2.log): the ring is bit-exact but SLOWER, 15.2 us vs
                             //    12.4 us per 1 Mi boards with 2 to 5 stages alike — the per-thread loads were never
                             //    the limiter (the kernel is issue-bound) and the ring's barrier traffic adds ~40
                             //    instructions per board-warp.  Kept as a tested build variant, not shipped.
#ifndef G2048_STAGES
#define G2048_STAGES 4
#endif
#ifndef G2048_NOWAIT         // EXPERIMENT ONLY: the step kernel skips griddepcontrol.wait — no ordering against the previous
#define G2048_NOWAIT 0       // launch at all (wrong in general); measures what a finer-grained dependency could gain
#endif
#ifndef G2048_CHAIN_NO_LATE_WAIT   // EXPERIMENT ONLY: chained launches never wait for the previous grid (breaks the stream
#define G2048_CHAIN_NO_LATE_WAIT 0 // order for whatever follows them); measures what the wait at their end costs
#endif
#ifndef G2048_LUT_WAIT_EARLY // 1: wait for the fresh-board table copy before the first board is loaded (round 2's order)
#define G2048_LUT_WAIT_EARLY 0
#endif
#ifndef G2048_LUT_GLOBAL     // EXPERIMENT: the 16 KB fresh-board table is read from global memory (L1-cached) instead of a
#define G2048_LUT_GLOBAL 0   // per-CTA shared-memory copy
#endif
constexpr int kCtasPerSm = G2048_CTAS_PER_SM;
// The step kernels' own persistent shape.  One 1024-thread CTA per SM measured 1 % faster than two of 512
// (12.23 vs 12.37 us per 1 Mi boards): with two CTAs the hardware scheduler favours the older one and the SM
// spends the last third of the launch on the younger one's 16 warps alone (scripts/micro/timeline.cu).
#ifndef G2048_STEP_THREADS            // (the TMA-ring build variant keeps 512 x 2: its ring is static shared memory)
#define G2048_STEP_THREADS (G2048_TMA ? 512 : 1024)
#endif
#ifndef G2048_STEP_CTAS_PER_SM
#define G2048_STEP_CTAS_PER_SM (G2048_TMA ? 2 : 1)
#endif
constexpr int kStepThreads = G2048_STEP_THREADS, kStepCtasPerSm = G2048_STEP_CTAS_PER_SM;
#ifndef G2048_STEP_MAXREG       // experiments: cap the step kernels' registers (0 = what the launch bounds allow: 64)
#define G2048_STEP_MAXREG 0
#endif
#if G2048_STEP_MAXREG
#define G2048_STEP_BOUNDS __maxnreg__(G2048_STEP_MAXREG)
#else
#define G2048_STEP_BOUNDS __launch_bounds__(kStepThreads, kStepCtasPerSm)
#endif
static_assert(kThreads == G2048_THREADS, "g2048_internal.h and g2048.cu disagree on the CTA size");

// The current device of the calling thread.  A launch list (g2048_step_list) issues thousands of launches on one
// device: it pins the answer for its duration instead of asking the runtime twice per launch.
static thread_local int tl_device_hint = -1;
static cudaError_t current_device(int* dev) {
  if (tl_device_hint >= 0) { *dev = tl_device_hint; return cudaSuccess; }
  return cudaGetDevice(dev);
}
struct DeviceHint {
  bool set;
  DeviceHint() : set(false) {
    int dev = -1;
    if (tl_device_hint < 0 && cudaGetDevice(&dev) == cudaSuccess) { tl_device_hint = dev; set = true; }
  }
  ~DeviceHint() { if (set) tl_device_hint = -1; }
};

static int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (current_device(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached = 148;
    cached_dev = dev;
  }
  return cached;
}
unsigned grid_for(uint64_t n) {
  const uint64_t need = (n + kThreads - 1) / kThreads;
  const uint64_t cap = (uint64_t)sm_count() * kCtasPerSm;
  return (unsigned)(need < cap ? need : cap);
}
// The streaming kernels (observation encoder, move, status, symmetries, ...) use few registers, so four of
// their CTAs fit an SM — and the block scheduler fills an SM to its residency limit before it moves on: a
// grid of two CTAs per SM would run on half the SMs.  Their persistent grid is therefore four CTAs per SM
// (observation encoder 84 -> 57 us per 1 Mi boards, profiles/r01_kernels.md).
unsigned grid_for_streaming(uint64_t n) {
  const uint64_t need = (n + kThreads - 1) / kThreads;
  const uint64_t cap = (uint64_t)sm_count() * 4u;
  return (unsigned)(need < cap ? need : cap);
}

// Launch shape of the step kernels.  A batch that fills the machine runs as one persistent wave of
// kStepCtasPerSm kStepThreads-wide CTAs per SM.  A smaller one is cut into 128-thread CTAs, and because the block
// scheduler packs CTAs onto an SM up to its residency limit before it moves on to the next SM (65,536 boards as
// 128 CTAs of 512 threads land on 64 SMs, two each, and leave 84 idle), the residency limit itself is set to
// the even share, ceil(CTAs / SMs), by padding every CTA with dynamic shared memory it never touches.
struct LaunchShape { unsigned grid, block; size_t pad_smem; };
#ifndef G2048_SMALL_BLOCK      // 0: choose the CTA width of a small batch from its size (below); else force it (experiments)
#define G2048_SMALL_BLOCK 0
#endif
static LaunchShape shape_for(uint64_t n, size_t static_smem, bool single_step) {
  const uint64_t sms = (uint64_t)sm_count();
  const uint64_t full_ctas = (n + kStepThreads - 1) / kStepThreads;
  if (full_ctas >= sms * kStepCtasPerSm) return LaunchShape{(unsigned)(sms * kStepCtasPerSm), (unsigned)kStepThreads, 0};
  // One board per thread.  The widest CTA that still gives every SM one: each CTA pays a prologue (a 16 KB table
  // copy, a barrier) and narrow CTAs multiply it — 131,072 boards run in 2.77 us as 256 CTAs of 512 threads and in
  // 3.47 us as 1024 CTAs of 128 (profiles/r02_variants.log) — while too few CTAs leave SMs idle (65,536 boards as
  // 64 CTAs of 1024 use 64 of 148 SMs).
  // (The multi-step kernel runs for many steps per launch: its prologue is amortised, it keeps 128-thread CTAs.)
  unsigned block = G2048_SMALL_BLOCK;
  if (block == 0) {
    block = single_step ? 1024u : 128u;
    while (block > 128u && (n + block - 1) / block < sms) block >>= 1;
  }
  const uint64_t ctas = (n + block - 1) / block;
  const uint64_t per_sm = (ctas + sms - 1) / sms;                          // 1 .. 8
  const size_t budget = (size_t)220 * 1024 / per_sm;                       // of the 227 KB an SM can carve out
  const size_t fixed = static_smem + 1024 + 256;                           // + the per-CTA reserve
  return LaunchShape{(unsigned)ctas, block, budget > fixed ? (budget - fixed) / 1024 * 1024 : 0};
}
constexpr int kMaxPadSmem = 208 * 1024;
// Kernel attributes are per device and are set ONCE PER PROCESS for every device (one bit per device ordinal in
// `seen`), under a lock: a second issuing thread (StepSchedule.run_threads, bench.py --issue-threads) must neither
// repeat the ~40 cudaFuncSetAttribute calls nor make them while the first thread is launching the same kernels
// (legal for the runtime, but compute-sanitizer's racecheck dies on it with plain launches of padded CTAs).
// Runs `set_attributes` the first time the current device is seen; every caller returns after they are set.
template <typename F> static void once_per_device(std::atomic<uint64_t>& seen, std::mutex& lock, F set_attributes) {
  int dev = 0;
  if (current_device(&dev) != cudaSuccess || dev < 0 || dev > 63) { set_attributes(); return; }
  const uint64_t bit = 1ull << dev;
  if (seen.load(std::memory_order_acquire) & bit) return;
  std::lock_guard<std::mutex> hold(lock);
  if (seen.load(std::memory_order_relaxed) & bit) return;
  set_attributes();
  seen.fetch_or(bit, std::memory_order_release);
}

// ------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------
constexpr uint32_t kFlagBumpCounter = 0x80000000u;   // internal: this launch advances *step_counter when it ends
constexpr uint32_t kFlagPrefetch = 0x40000000u;      // internal: L2-prefetch the first boards before griddepcontrol.wait
constexpr uint32_t kFlagChained = 0x20000000u;       // internal: G2048_FLAG_CHAINED — the launch waits per slice, not per grid
// Chain buffer (G2048StepArgs.chain, G2048_CHAIN_WORDS 64-bit words): one word per WARP of a chained launch shape
// (kChainSlices CTAs of kChainThreads threads at most) = the step index + 1 of the last launch whose warp there is done.
#ifndef G2048_CHAIN_THREADS          // launch shape of a step that carries a chain buffer (experiments: scripts/chain_variants.sh)
#define G2048_CHAIN_THREADS 256
#endif
#ifndef G2048_CHAIN_CTAS_PER_SM           // a chain whose launches follow each other directly (every launch waits for its
#define G2048_CHAIN_CTAS_PER_SM 4         // predecessor's warps): a wide launch, the fifth place on an SM is for the next one
#endif
#ifndef G2048_CHAIN_NARROW_CTAS_PER_SM    // G2048_FLAG_CHAIN_INTERLEAVED: launches of other chains in between — several
#define G2048_CHAIN_NARROW_CTAS_PER_SM 1  // launches share the machine side by side, each one CTA per SM
#endif
constexpr unsigned kChainThreads = G2048_CHAIN_THREADS, kChainCtasPerSm = G2048_CHAIN_CTAS_PER_SM,
                   kChainNarrowCtasPerSm = G2048_CHAIN_NARROW_CTAS_PER_SM, kChainNarrowMaxIters = 16u;
// (a call whose env ids cross a multiple of 2^32 is two launches: the second one uses the upper half of the buffer)
constexpr unsigned kChainLaunchWords = G2048_CHAIN_WORDS / 2u;
constexpr unsigned kChainSlices = kChainLaunchWords / (kChainThreads / 32u);     // CTAs a launch has words for
static_assert(kChainThreads % 32u == 0u && kChainThreads <= 1024u && kChainSlices >= 148u * kChainCtasPerSm, "chain shape");
#ifndef G2048_PREFETCH_MIN_N   // the prologue prefetch pays from this batch size on (3.47 -> 3.16 us WITHOUT it at 131,072
#define G2048_PREFETCH_MIN_N 200000   // boards, 4.74 -> 4.24 us with it at 262,144, 11.5 -> 11.0 us at 1 Mi; profiles/r02_variants.log)
#endif

struct StepParams {
  const uint4* boards;
  uint4* boards_out;            // == boards for an in-place step
  const uint8_t* actions;
  float* rewards;
  uint8_t* dones;
  uint8_t* illegal;
  uint8_t* highest_exp;
  uint8_t* legal_mask;
  uint4* terminal_boards;
  uint32_t* ep_score;
  uint32_t* ep_len;
  uint32_t* final_score;
  uint32_t* final_len;
  float* ep_return;
  float* final_return;
  uint2* boards_nibble;         // [n] the board handed back, 4 bits per cell (cell c at bits 4c..4c+3 of the 64-bit word)
  uint32_t* nibble_overflow;    // counter of boards whose exponents do not fit 4 bits (a tile >= 65,536)
  const uint4* forced_draws;
  uint64_t* step_counter;       // [0] step index, [1] CTA arrival ticket (0 between launches)
  unsigned long long* chain;    // fine-grained dependency between consecutive launches over the same boards (see the kernel)
  unsigned long long chain_tag; // step index + 1: what a warp publishes when it is done; a chained warp waits for chain_tag - 1
  uint32_t n;                   // < 2^32 - 256 (checked by the host)
  uint32_t env_lo;              // low half of the env id of board 0; the launch never crosses 2^32
  uint32_t env_hi;              // high half, the same for every board
  uint64_t seed;                // only read when step_counter is set (the key is then derived on the device)
  StreamKeys keys;              // Philox2x32 round keys of stream_key(seed, step_index, env_hi, TAG_STEP)
  StreamKeys policy_keys;       // the same for TAG_POLICY (kernels that draw the actions themselves)
  uint8_t* actions_out;         // POLICY kernels: where the drawn action goes (the caller's `actions` array)
  float illegal_move_reward;
  uint32_t max_tile_exp;
  uint32_t flags;
};

// The 1024 fresh boards of two_tile_board() (g2048_device.cuh), built at compile time: the step kernel
// brings the table into shared memory with ONE bulk async copy (TMA) per CTA, issued before
// griddepcontrol.wait — it is constant data, so the copy overlaps the previous launch's tail — instead of
// spending ~60 instructions per thread rebuilding it every launch.
struct alignas(128) PairLut { Board4 e[1024]; };
constexpr PairLut make_pair_lut() {
  PairLut t{};
  for (uint32_t entry = 0; entry < 1024u; ++entry) t.e[entry] = two_tile_board(entry);
  return t;
}
__device__ const PairLut g_pair_lut = make_pair_lut();

// Game2048Env.step (:76-100) for n boards.  OUT = 0 is the lean variant used when none of the
// optional outputs/inputs is requested (boards, actions, rewards, dones only); see O_* below.
// The 32 one-tile boards fresh_board() ORs together, built once per CTA in shared memory.
__device__ __forceinline__ const Board4* make_reset_lut(Board4* s_lut) {
  if (threadIdx.x < 32) s_lut[threadIdx.x] = one_tile_board(threadIdx.x);
  __syncthreads();
  return s_lut;
}

#ifndef G2048_LD_HINT        // 0: plain ld.global; 1: L1::no_allocate (boards are read once; measured -2 %); 2: ld.global.cs
#define G2048_LD_HINT 1
#endif
__device__ __forceinline__ uint4 load_board(const uint4* ptr) {
#if G2048_LD_HINT == 1
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
#elif G2048_LD_HINT == 2
  return __ldcs(ptr);
#else
  return *ptr;
#endif
}

// One board, already rotated into its move frame (a,b,c,d), through Game2048Env.step and out to memory.
// COUNTER: the step index is read from device memory (CUDA-graph replay), so the generator's key
// and counter word 1 are registers (dev_key, dev_idx_lo) instead of kernel-parameter constants.
#ifndef G2048_PAIR_LUT       // 1: the step kernel's reset path reads whole fresh boards from a 16 KB shared-memory
#define G2048_PAIR_LUT 1     //    table (two_tile_board) instead of OR-ing two entries of the 32-entry one-tile table
#endif
// Optional outputs/inputs of a step as a COMPILE-TIME set (template parameter OUT of the step kernel): the output
// sets the library's own callers use are kernels of their own — no pointer tests, no dead code, registers only for
// what is written — and any other combination runs the O_GENERIC kernel, which tests every pointer at run time.
//   0                                  lean: boards, actions, rewards, dones            (BASELINE configs 2, 3)
//   O_MASK                             + legal mask                                     (BASELINE config 4)
//   O_EPRUN                            + running episode score/length                   (g2048_env_step_host, `e2e`)
//   O_EPRUN|O_NIBBLE                   + the board packed 4 bits per cell               (the same, compact host format)
//   O_ILLEGAL|O_HIGHEST|O_MASK         the evaluator (train.py:122-214)
//   kOutAll                            everything SB3's Monitor + DummyVecEnv report    (Game2048VecEnv)
constexpr uint32_t O_ILLEGAL = 1u, O_HIGHEST = 2u, O_MASK = 4u, O_TERMINAL = 8u, O_EPRUN = 16u, O_EPFINAL = 32u,
                   O_EPRET = 64u, O_FORCED = 128u, O_NIBBLE = 256u, O_GENERIC = 0x80000000u;
constexpr uint32_t kOutEval = O_ILLEGAL | O_HIGHEST | O_MASK;
constexpr uint32_t kOutAll = O_ILLEGAL | O_HIGHEST | O_MASK | O_TERMINAL | O_EPRUN | O_EPFINAL | O_EPRET;
template <uint32_t OUT, uint32_t BIT> __device__ __forceinline__ bool has(const void* ptr) {
  if constexpr ((OUT & O_GENERIC) != 0u) return ptr != nullptr;
  else return (OUT & BIT) != 0u;
}

// The running episode statistics of a board, loaded AHEAD of the step (with the board, one loop iteration early):
// read where they are used they cost an exposed HBM round trip per board (ncu: long-scoreboard stalls 0.7 -> 4.3
// per issue cycle, 15 -> 17.8 us per 1 Mi boards for the O_EPRUN kernel).
struct EpIn { uint32_t score, len; float ret; };
template <uint32_t OUT>
__device__ __forceinline__ EpIn load_episode(const StepParams& p, uint32_t i) {
  EpIn e{0u, 0u, 0.f};
  if constexpr (OUT != 0u) {
    if (has<OUT, O_EPRUN>(p.ep_score)) e.score = p.ep_score[i];
    if (has<OUT, O_EPRUN>(p.ep_len)) e.len = p.ep_len[i];
    if (has<OUT, O_EPRET>(p.ep_return)) e.ret = p.ep_return[i];
  }
  return e;
}

// The finishing half of board i (spawn, score, isend, auto-reset) and its stores; `m` is what the
// move half (move_oriented) left, `ep` the board's running episode statistics before the step.
template <uint32_t OUT, bool COUNTER>
__device__ __forceinline__ void finish_and_store(const StepParams& p, const Board4* lut, uint32_t i, const Moved& m,
                                                 uint32_t dev_key, uint32_t dev_idx_lo, bool auto_reset, const EpIn ep) {
  Words w;
  if (has<OUT, O_FORCED>(p.forced_draws)) {
    const uint4 f = p.forced_draws[i];
    w = Words{f.x, f.y, f.z, f.w};
  } else if (COUNTER) {
    w = words_from_pair(philox2x32_10(p.env_lo + i, dev_idx_lo, dev_key));
  } else {
    w = words_from_pair(philox2x32_10_keys(p.env_lo + i, p.keys));
  }
  uint4 bd;
#ifndef G2048_FORCE_PLAIN     // experiment: max_tile = None and auto-reset as compile-time facts (what a PLAIN flavour would save)
#define G2048_FORCE_PLAIN 0
#endif
  const StepOut o = finish_step<(G2048_PAIR_LUT && !G2048_TMA)>(lut, m, w, G2048_FORCE_PLAIN ? 0u : p.max_tile_exp, has<OUT, O_HIGHEST>(p.highest_exp),
                                                    auto_reset, bd.x, bd.y, bd.z, bd.w);
  const float reward = o.legal ? o.score : p.illegal_move_reward;                        // :90 / :95
  p.boards_out[i] = bd;
  p.rewards[i] = reward;
  p.dones[i] = o.done ? 1 : 0;
  if constexpr (OUT != 0u) {
    if (has<OUT, O_ILLEGAL>(p.illegal)) p.illegal[i] = o.legal ? 0 : 1;                  // :82, :93
    if (has<OUT, O_HIGHEST>(p.highest_exp)) p.highest_exp[i] = (uint8_t)o.highest;       // :97
    // (in a specialised kernel a group bit stands for all pointers of the group; the generic one tests each)
    const bool h_es = has<OUT, O_EPRUN>(p.ep_score), h_el = has<OUT, O_EPRUN>(p.ep_len);
    const bool h_er = has<OUT, O_EPRET>(p.ep_return);
    uint32_t es = 0, el = 0;
    float er = 0.f;
    if (h_es) es = ep.score + (uint32_t)o.score;                                         // :86
    if (h_el) el = ep.len + 1u;
    if (h_er) er = ep.ret + reward;                      // what SB3's Monitor sums: the rewards the agent saw
    if (o.done) {
      if (has<OUT, O_TERMINAL>(p.terminal_boards)) p.terminal_boards[i] = make_uint4(o.t0, o.t1, o.t2, o.t3);
      if (has<OUT, O_EPFINAL>(p.final_score)) p.final_score[i] = es;
      if (has<OUT, O_EPFINAL>(p.final_len)) p.final_len[i] = el;
      if (has<OUT, O_EPRET>(p.final_return)) p.final_return[i] = er;
      if (auto_reset) { es = el = 0u; er = 0.f; }
    }
    if (h_es) p.ep_score[i] = es;
    if (h_el) p.ep_len[i] = el;
    if (h_er) p.ep_return[i] = er;
    if (has<OUT, O_MASK>(p.legal_mask)) p.legal_mask[i] = (uint8_t)legal_mask(bd.x, bd.y, bd.z, bd.w);
    if (has<OUT, O_NIBBLE>(p.boards_nibble)) {
      // 16 exponents -> 16 nibbles (the classic 64-bit 2048 board): per row, the low nibbles of the four bytes
      auto pack_row = [](uint32_t r) {
        uint32_t x = r & 0x0F0F0F0Fu;
        x = (x | (x >> 4)) & 0x00FF00FFu;
        return (x | (x >> 8)) & 0xFFFFu;
      };
      p.boards_nibble[i] = make_uint2(pack_row(bd.x) | (pack_row(bd.y) << 16), pack_row(bd.z) | (pack_row(bd.w) << 16));
      if ((((bd.x | bd.y) | (bd.z | bd.w)) & 0xF0F0F0F0u) != 0u && p.nibble_overflow) atomicAdd(p.nibble_overflow, 1u);
    }
  }
}

template <uint32_t OUT, bool COUNTER>
__device__ __forceinline__ void step_and_store(const StepParams& p, const Board4* lut, uint32_t i, uint32_t a,
                                               uint32_t b, uint32_t c, uint32_t d, const Sel4 so,
                                               uint32_t dev_key, uint32_t dev_idx_lo, bool auto_reset, const EpIn ep) {
  finish_and_store<OUT, COUNTER>(p, lut, i, move_oriented(a, b, c, d, so), dev_key, dev_idx_lo, auto_reset, ep);
}

#ifndef G2048_PIPELINE       // 1: software-pipelined loop — the move half of board j+1 and the finishing half of
#define G2048_PIPELINE 0     //    board j share one loop body (one straight-line block the scheduler interleaves).
#endif                       //    It was 2 % ahead of the plain loop when it was written (512x2 CTAs, 32-entry reset
                             //    table); with the fresh-board table and one 1024-thread CTA per SM the plain loop is
                             //    2 % ahead (11.99 vs 12.22 us, profiles/r01_variants_v2.log), so the plain loop ships.
// Tried on top of this loop and dropped, both bit-exact and both slower because the kernel is issue-bound and every
// extra instruction costs more than the memory system gives back (same log): warps claiming 32-board tiles from a
// per-CTA shared-memory counter so that no warp runs out of work early (13.7 us vs 12.4 us: +30 instructions per
// board-warp for the claim and the ragged-tile predicates), and an L2 bulk prefetch (cp.async.bulk.prefetch.L2) two
// iterations ahead (13.2 us vs 12.4 us).
#ifndef G2048_SEL_SMEM       // 1: orientation selectors from shared memory (LDS.128, conflict-free) instead of
#define G2048_SEL_SMEM 1     //    action-indexed constant memory (a divergent LDC is replayed per distinct address)
#endif
#ifndef G2048_PTR_INC        // 1: walk the input arrays with pointers kept in registers instead of re-reading the
#define G2048_PTR_INC 0      //    kernel parameters (LDC) and re-deriving the addresses every iteration.  Measured:
#endif                       //    smem selectors -1.2 %, pointer walk -0.7 %, both together -0.6 % -> selectors only

[[maybe_unused]] constexpr int kStages = G2048_STAGES;

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Bulk async copy global -> shared (TMA, 1-D): `bytes` a multiple of 16, both addresses 16-byte aligned;
// completion is signalled as transaction bytes on `bar`.
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// POLICY 0: the actions are read from p.actions.  1 / 2: the kernel plays g2048_sample_actions' uniform / random-legal
// policy itself — the action is drawn from the policy-tag stream at this step's index, for the legal policy among
// the legal moves p.legal_mask holds for the board (the mask the previous step wrote) — and writes it to
// p.actions_out: BASELINE config 4's "sample a legal action, step, return the new mask" is ONE launch.
template <uint32_t OUT, bool COUNTER, int POLICY = 0>
__global__ void G2048_STEP_BOUNDS g2048_step_kernel(const StepParams p) {
  static_assert(POLICY == 0 || (!G2048_TMA && !G2048_PIPELINE), "the policy kernels exist for the plain loop only");
  static_assert(POLICY != 2 || (OUT & (O_MASK | O_GENERIC)) != 0u, "the random-legal policy reads and writes the legal mask");
#if G2048_LUT_GLOBAL
  // (experiment) the fresh-board table is read where it lies, through the L1: no per-CTA copy, no barrier
#elif G2048_PAIR_LUT && !G2048_TMA
  __shared__ alignas(128) Board4 s_lut[1024];
  __shared__ alignas(8) uint64_t s_lut_bar;
  if (threadIdx.x == 0) {
    mbar_init(&s_lut_bar, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_lut_bar, (uint32_t)sizeof(PairLut));
    bulk_load(s_lut, &g_pair_lut, (uint32_t)sizeof(PairLut), &s_lut_bar);
  }
#else
  __shared__ Board4 s_lut[32];
#endif
  __shared__ Sel4 s_sel[8];     // [action] = kOrientIn, [4 + action] = kOrientOut
#if G2048_TMA
  __shared__ alignas(128) uint4 s_boards[kStages][kStepThreads];
  __shared__ alignas(16) uint8_t s_actions[kStages][kStepThreads];
  __shared__ alignas(8) uint64_t s_full[kStages], s_empty[kStages];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kStages; ++k) { mbar_init(&s_full[k], 1u); mbar_init(&s_empty[k], kStepThreads / 32u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#endif
  // ---- chained launches (G2048StepArgs.chain) ----------------------------------------------------------------
  // A step over boards that the PREVIOUS step launch wrote depends on that launch board by board, not grid by grid:
  // thread t of CTA b owns the same boards in every launch of the same shape.  With a chain buffer every warp
  // publishes "step index + 1" in its own word when its last store is out (release at GPU scope), and a launch that
  // carries kFlagChained does not wait for the previous GRID (griddepcontrol.wait) but, warp by warp, for that word
  // to show the previous step (acquire).  The launches are programmatic dependent launches whose CTAs signal
  // launch_dependents at entry, so the hardware starts a CTA of the next launch as soon as an SM has room for it:
  // its prologue runs in the shadow of the CTAs still at work there, and there is no grid-wide drain, release
  // and ramp between two steps (2.35 us of an 11 us launch over 1 Mi boards).  No deadlock: a CTA of launch k+1
  // exists only after every CTA of launch k has started.  A launch WITHOUT the flag (the head of a chain, or a
  // step after foreign work touched its inputs) keeps griddepcontrol.wait; it still publishes.
  const bool chain = p.chain != nullptr;
  const bool chained = chain && (p.flags & kFlagChained) != 0u;
  // (the first look at the warp's word is taken right here, so that its round trip to the L2 overlaps the prologue)
  const unsigned long long* const chain_word = p.chain + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  unsigned long long chain_seen = 0ull;
  if (chained && (threadIdx.x & 31u) == 0u)
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(chain_seen) : "l"(chain_word) : "memory");
#if G2048_PDL
  // Let the next launch in the stream start its ramp as soon as our CTAs retire; everything
  // before griddepcontrol.wait touches no global memory a previous launch could have written (only the
  // constant fresh-board table), so it overlaps the previous kernel.
  asm volatile("griddepcontrol.launch_dependents;");
#endif
  if (threadIdx.x >= 32 && threadIdx.x < 40) {
    const uint32_t k = threadIdx.x - 32;
    s_sel[k] = (k < 4) ? kOrientIn[k] : kOrientOut[k - 4];
  }
#if G2048_LUT_GLOBAL
  __syncthreads();                          // s_sel written
  const Board4* lut = g_pair_lut.e;
#elif G2048_PAIR_LUT && !G2048_TMA
  __syncthreads();                          // s_sel written, s_lut_bar initialised
  const Board4* lut = s_lut;
#else
  const Board4* lut = make_reset_lut(s_lut);
#endif
  const bool auto_reset = G2048_FORCE_PLAIN ? true : (p.flags & G2048_FLAG_AUTO_RESET) != 0u;
  const uint32_t n = p.n;
#if !G2048_TMA
  const uint32_t stride = gridDim.x * blockDim.x;               // the CTA width is chosen at launch (shape_for)
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
#endif
#if G2048_PRE_PREFETCH && !G2048_TMA
  // Cold start: after griddepcontrol.wait every warp's first board is an HBM round trip (~1 us of a ~12 us launch)
  // with nothing to compute meanwhile.  An L2 prefetch has no architectural effect (the L2 is the point of
  // coherence: a line the previous launch is still writing is simply already there), so the first lines can be
  // requested while the previous launch drains: one lane per 128-byte line (8 boards), one per warp for the actions.
  if (i < n && (p.flags & kFlagPrefetch)) {
    if ((threadIdx.x & 7u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.boards + i));
    if (POLICY != 1 && (threadIdx.x & 31u) == 0u)
      asm volatile("prefetch.global.L2 [%0];" ::"l"((POLICY == 2 ? p.legal_mask : p.actions) + i));
  }
#endif
#if G2048_PDL && !G2048_NOWAIT
  if (chained) {
    if ((threadIdx.x & 31u) == 0u) {
      const unsigned long long want = p.chain_tag - 1ull;      // what this warp's predecessor published
      uint32_t polls = 0u;
      while (chain_seen != want) {
        // back off: a predecessor that is merely slow (a debugger, a sanitizer, a time-sliced GPU) must not be
        // mistaken for one that never ran — the trap comes after more than 30 s of waiting
        if (++polls > 16u) __nanosleep(polls > 4096u ? 4000u : 64u);
        if (polls > (1u << 23)) __trap();                      // the caller chained to a launch that never ran
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(chain_seen) : "l"(chain_word) : "memory");
      }
    }
    __syncwarp();
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
#endif
#if G2048_PAIR_LUT && !G2048_TMA && !G2048_LUT_GLOBAL && (G2048_PIPELINE || G2048_LUT_WAIT_EARLY)
  mbar_wait(&s_lut_bar, 0u);                // the table copy was started in the prologue; long done by now
#endif
  uint64_t counter_value = 0;
  uint32_t dev_key = 0u, dev_idx_lo = 0u;
  [[maybe_unused]] uint32_t dev_policy_key = 0u;
  if (COUNTER) {                            // device-side step index (CUDA-graph replay)
    counter_value = p.step_counter[0];
    dev_key = stream_key(p.seed, counter_value, (uint64_t)p.env_hi << 32, TAG_STEP);
    dev_idx_lo = (uint32_t)counter_value;
    if (POLICY) dev_policy_key = stream_key(p.seed, counter_value, (uint64_t)p.env_hi << 32, TAG_POLICY);
    __syncthreads();                        // every thread of the CTA has read the index before thread 0 can arrive
  }
#if !G2048_TMA
  if (i >= n && p.chain == nullptr) return; // (a chained warp stays whole: it publishes after its last store)
#endif
  // Grid-stride loop, software-pipelined one board ahead.  orient() consumes the loaded board
  // right away, so the next board is prefetched into the same registers (no rotation copies)
  // a whole iteration before it is used.  (Unrolling by two was measured slower: the doubled
  // body no longer fits the L0 instruction cache; so were two boards per thread.)
#if G2048_TMA
  // Staged loop.  The CTA walks tiles of kStepThreads consecutive boards (tile t = blockIdx.x + k * gridDim.x).  One
  // elected thread keeps kStages - 1 tiles in flight: per tile two bulk async copies (TMA, cp.async.bulk) bring
  // the 16-byte boards and the action bytes into a shared-memory ring and complete on the stage's `full`
  // mbarrier; every thread waits on that barrier, takes its board and action with one LDS.128 and one LDS.U8,
  // and each warp then arrives on the stage's `empty` barrier, which the producer waits on before it refills
  // the stage.  Loads are thus decoupled from the registers and run several iterations ahead of the compute.
  const uint32_t tid = threadIdx.x;
  const uint32_t n_tiles = (n + kStepThreads - 1u) / kStepThreads;
  const bool actions_by_tma = (reinterpret_cast<uintptr_t>(p.actions) & 15u) == 0u;
  auto tile_count = [&](uint32_t t) { const uint32_t left = n - t * kStepThreads; return left < (uint32_t)kStepThreads ? left : (uint32_t)kStepThreads; };
  auto produce = [&](uint32_t t, uint32_t stage) {      // one thread: arm the barrier, start the copies of tile t
    const uint32_t cnt = tile_count(t);
    const bool act_tma = actions_by_tma && (cnt & 15u) == 0u;
    mbar_expect_tx(&s_full[stage], cnt * 16u + (act_tma ? cnt : 0u));
    bulk_load(&s_boards[stage][0], p.boards + (size_t)t * kStepThreads, cnt * 16u, &s_full[stage]);
    if (act_tma) bulk_load(&s_actions[stage][0], p.actions + (size_t)t * kStepThreads, cnt, &s_full[stage]);
  };
  if (tid == 0) {
#pragma unroll
    for (uint32_t k = 0; k < (uint32_t)kStages - 1u; ++k) {
      const uint32_t t = blockIdx.x + k * gridDim.x;
      if (t < n_tiles) produce(t, k);
    }
  }
  uint32_t stage = 0, parity = 0;                 // ring position of iteration k and its use count's parity
  uint32_t fill_stage = kStages - 1u, fill_parity = 1u;   // where iteration k's refill (tile k + kStages - 1) goes
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const uint32_t cnt = tile_count(t);
    const bool act_tma = actions_by_tma && (cnt & 15u) == 0u;
    const uint32_t i = t * kStepThreads + tid;
    const bool valid = tid < cnt;
    if (tid == 0) {
      // Refill the stage the CTA consumed in the previous iteration (all warps have normally left it).
      const uint32_t t_fill = t + ((uint32_t)kStages - 1u) * gridDim.x;
      if (t_fill < n_tiles && t_fill > t) {
        mbar_wait(&s_empty[fill_stage], fill_parity);
        produce(t_fill, fill_stage);
      }
    }
    mbar_wait(&s_full[stage], parity);
    const uint4 bd = s_boards[stage][tid];
    uint32_t action = s_actions[stage][tid];
    if (!act_tma) action = valid ? p.actions[i] : 0u;
    __syncwarp();
    if ((tid & 31u) == 0u) mbar_arrive(&s_empty[stage]);
    const uint32_t act = action & 3u;
    uint32_t a, b, c, d;
    orient(s_sel[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
    const Sel4 so = s_sel[4u + act];
    if (valid) step_and_store<OUT, COUNTER>(p, lut, i, a, b, c, d, so, dev_key, dev_idx_lo, auto_reset, load_episode<OUT>(p, i));
    if (++stage == (uint32_t)kStages) { stage = 0; parity ^= 1u; }
    if (++fill_stage == (uint32_t)kStages) { fill_stage = 0; fill_parity ^= 1u; }
  }
#elif G2048_PIPELINE
  // Software-pipelined grid-stride loop.  Iteration j runs the MOVE half of board j+1 (orient, slide+merge,
  // un-orient: PRMT/LOP3, the ALU pipe) and the FINISHING half of board j (Philox, spawn, float score, isend,
  // reset, stores: IMAD-heavy, the FMA pipe) as one straight-line block; the two halves are independent, so the
  // scheduler interleaves them and both pipes stay busy.  Seven registers (Moved) carry a board from one half
  // to the other.  The board after next is prefetched into the registers orient() has just freed.
  {
    uint4 bd = load_board(p.boards + i);
    uint32_t action = p.actions[i];
    uint32_t i_next = i + stride;
    bool more = G2048_PERSISTENT && i_next < n && i_next > i;
    Moved m;
    {
      const uint32_t act = action & 3u;
      uint32_t a, b, c, d;
      orient(s_sel[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
      const Sel4 so = s_sel[4u + act];
      if (more) { bd = load_board(p.boards + i_next); action = p.actions[i_next]; }
      m = move_oriented(a, b, c, d, so);
    }
    while (more) {
      const uint32_t i_cur = i_next;
      i_next = i_cur + stride;
      const bool more_next = i_next < n && i_next > i_cur;
      const uint32_t act = action & 3u;
      uint32_t a, b, c, d;
      orient(s_sel[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
      const Sel4 so = s_sel[4u + act];
      if (more_next) { bd = load_board(p.boards + i_next); action = p.actions[i_next]; }
      const Moved m_next = move_oriented(a, b, c, d, so);
      finish_and_store<OUT, COUNTER>(p, lut, i, m, dev_key, dev_idx_lo, auto_reset, load_episode<OUT>(p, i));
      m = m_next;
      i = i_cur;
      more = more_next;
    }
    finish_and_store<OUT, COUNTER>(p, lut, i, m, dev_key, dev_idx_lo, auto_reset, load_episode<OUT>(p, i));
  }
#else
  if (i < n) {
  const uint4* pb = p.boards + i;
  // the byte prefetched next to the board: the action, or (random-legal policy) the board's legal mask
  const uint8_t* const side = POLICY == 2 ? p.legal_mask : p.actions;
  const uint8_t* pa = side + i;
  uint4 bd = load_board(pb);
  uint32_t action = POLICY == 1 ? 15u : *pa;
  EpIn ep_next = load_episode<OUT>(p, i);
#if G2048_PAIR_LUT && !G2048_LUT_GLOBAL && !G2048_LUT_WAIT_EARLY
  // The fresh-board table (its copy was started in the prologue) is first read at the end of a step: wait for it
  // with the first board's load already in flight.  (A chained CTA starts working as soon as it is resident: its
  // table copy is not "long done" as it is behind griddepcontrol.wait.)
  mbar_wait(&s_lut_bar, 0u);
#endif
#if G2048_STAGGER
  // Experiment: the eight warps of an SM sub-partition run the same instruction stream in near lockstep — all in
  // the ALU-heavy move, then all in the FMA-heavy Philox/spawn — so the two pipes take turns idling.  Start them
  // G2048_STAGGER ns apart (their first loads are in flight meanwhile).
  __nanosleep((threadIdx.x >> 7) * G2048_STAGGER);
#endif
  while (true) {
    const uint32_t i_next = i + stride;
    const bool more = G2048_PERSISTENT && i_next < n && i_next > i;
    uint32_t act;
    if (POLICY) {
      const Pair pw = COUNTER ? philox2x32_10(p.env_lo + i, dev_idx_lo, dev_policy_key)
                              : philox2x32_10_keys(p.env_lo + i, p.policy_keys);
      act = pick_action(action, pw.x0);
      p.actions_out[i] = (uint8_t)act;
    } else {
      act = action & 3u;
    }
    uint32_t a, b, c, d;
#if G2048_SEL_SMEM
    orient(s_sel[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
    const Sel4 so = s_sel[4u + act];
#else
    orient(kOrientIn[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
    const Sel4 so = kOrientOut[act];
#endif
    const EpIn ep = ep_next;
#if G2048_PREFETCH
    if (more) {
      ep_next = load_episode<OUT>(p, i_next);
#if G2048_PTR_INC
      pb += stride; pa += stride;
      bd = load_board(pb);
      if (POLICY != 1) action = *pa;
#else
      bd = load_board(p.boards + i_next);
      if (POLICY != 1) action = side[i_next];
#endif
    }
#endif
    step_and_store<OUT, COUNTER>(p, lut, i, a, b, c, d, so, dev_key, dev_idx_lo, auto_reset, ep);
    if (!more) break;
#if !G2048_PREFETCH
    bd = load_board(p.boards + i_next);
    if (POLICY != 1) action = side[i_next];
    ep_next = load_episode<OUT>(p, i_next);
#endif
    i = i_next;
  }
  }
  if (p.chain != nullptr) {                  // (everything is re-derived from the parameters here: nothing of the chain
    // Publish: the warp's stores are ordered before lane 0's release store by the warp barrier, and the release
    // makes them visible at GPU scope before the word.                         stays in registers through the loop)
    __syncwarp();
    if ((threadIdx.x & 31u) == 0u)
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p.chain + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5))),
                   "l"(p.chain_tag) : "memory");
#if G2048_PDL && !G2048_NOWAIT && !G2048_CHAIN_NO_LATE_WAIT
    // A chained launch did not wait for the previous grid before its work — it waits for it now, before it ENDS.
    // Stream order is transitive only through kernels that wait: a kernel behind this one (a plain step of another
    // env set, any programmatic dependent launch) waits for THIS grid to complete and must be able to conclude that
    // everything issued before it has completed too.  The launch in front of us started earlier and is normally
    // done by now; the words above are already published, so no successor is held up.
    if (p.flags & kFlagChained) asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  }
#endif
  // The last CTA to arrive advances the device-side step index: by then every CTA has read it.
  // (Thread 0 of a CTA always owns a board, so it never took the early exit above.)
  if (COUNTER && (p.flags & kFlagBumpCounter) && threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(p.step_counter + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
      *ticket = 0u;
      p.step_counter[0] = counter_value + 1ull;
    }
  }
}

// ------------------------------------------------------------------------------------
// Several steps per launch (g2048_step_many): the same step, applied n_steps times to boards
// that stay in registers in between.  For open-loop action sequences — pre-generated random
// actions, replays of recorded games — the board traffic (32 of the 38 bytes of a step) and the
// per-step launch disappear; per step a thread reads one action byte and writes reward and done.
// Results are bit-identical to n_steps calls of g2048_step (step k draws with index step_index + k).
// ------------------------------------------------------------------------------------
struct ManyParams {
  uint4* boards;                // [n] in/out
  const uint8_t* actions;       // [n_steps][row_stride]
  float* rewards;               // [n_steps][row_stride]
  uint8_t* dones;               // [n_steps][row_stride]
  uint8_t* illegal;             // nullable, [n_steps][row_stride]
  uint4* boards_traj;           // nullable, [n_steps][row_stride]: the board handed back after every step
  uint8_t* actions_out;         // nullable, [n_steps][row_stride]: the actions a device policy chose
  uint8_t* legal_mask;          // nullable, [n_steps][row_stride]: legal moves on the board handed back
  uint32_t n;                   // boards of this launch
  uint32_t n_steps;
  uint64_t row_stride;          // boards per step row of the [n_steps][...] arrays (>= n when the call was sliced)
  uint32_t env_lo;              // low half of board 0's env id (the launch never crosses 2^32 ids)
  uint32_t idx_lo;              // low half of the first step's index (nor 2^32 indices)
  uint32_t key;                 // stream_key(seed, step_index, env id, TAG_STEP)
  StreamKeys keys;              // k[1..9] of that key (k[0] is derived per step)
  uint32_t policy_key;          // the same for TAG_POLICY (device policies)
  StreamKeys policy_keys;
  float illegal_move_reward;
  uint32_t max_tile_exp;
  uint32_t flags;
};

// POLICY 0: the actions are given.  1 / 2: the kernel plays g2048_sample_actions' uniform / random-legal policy
// itself — the action of step k is drawn from the policy-tag stream at index step_index + k, for the legal policy
// among the legal moves of the board the previous step handed back — so a whole random rollout is one launch.
template <bool EXTRAS, int POLICY>
__global__ void G2048_STEP_BOUNDS g2048_step_many_kernel(const ManyParams p) {
  __shared__ alignas(128) Board4 s_lut[1024];
  __shared__ alignas(8) uint64_t s_lut_bar;
  __shared__ Sel4 s_sel[8];
  if (threadIdx.x == 0) {
    mbar_init(&s_lut_bar, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_lut_bar, (uint32_t)sizeof(PairLut));
    bulk_load(s_lut, &g_pair_lut, (uint32_t)sizeof(PairLut), &s_lut_bar);
  }
#if G2048_PDL
  asm volatile("griddepcontrol.launch_dependents;");
#endif
  if (threadIdx.x >= 32 && threadIdx.x < 40) {
    const uint32_t k = threadIdx.x - 32;
    s_sel[k] = (k < 4) ? kOrientIn[k] : kOrientOut[k - 4];
  }
  __syncthreads();
#if G2048_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  mbar_wait(&s_lut_bar, 0u);
  const bool auto_reset = (p.flags & G2048_FLAG_AUTO_RESET) != 0u;
  const uint32_t n = p.n, stride = gridDim.x * blockDim.x;       // the CTA size is chosen at launch (see launch_step_many)
  // (i_next > i: n may be as large as 2^32 - 257, so i + stride can wrap — the same guard as in g2048_step_kernel)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, i_next; i < n; i = i_next > i ? i_next : n) {
    i_next = i + stride;
    uint4 bd = load_board(p.boards + i);
    size_t off = i;                                   // element (k, i) of the per-step arrays
    uint32_t action = POLICY ? 0u : p.actions[off];
    uint32_t mask = POLICY == 2 ? legal_mask(bd.x, bd.y, bd.z, bd.w) : 15u;
    for (uint32_t k = 0; k < p.n_steps; ++k, off += p.row_stride) {
      uint32_t act;
      if (POLICY) {
        act = pick_action(mask, philox2x32_10_keys(p.env_lo + i, p.policy_key ^ (p.idx_lo + k), p.policy_keys).x0);
        if (p.actions_out) p.actions_out[off] = (uint8_t)act;
      } else {
        act = action & 3u;
        if (k + 1u < p.n_steps) action = p.actions[off + p.row_stride];     // next step's action, one step ahead
      }
      uint32_t a, b, c, d;
      orient(s_sel[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
      const Moved m = move_oriented(a, b, c, d, s_sel[4u + act]);
      const Words w = words_from_pair(philox2x32_10_keys(p.env_lo + i, p.key ^ (p.idx_lo + k), p.keys));
      const StepOut o = finish_step<true>(s_lut, m, w, p.max_tile_exp, false, auto_reset, bd.x, bd.y, bd.z, bd.w);
      p.rewards[off] = o.legal ? o.score : p.illegal_move_reward;           // :90 / :95
      p.dones[off] = o.done ? 1 : 0;
      if (EXTRAS) {
        if (p.illegal) p.illegal[off] = o.legal ? 0 : 1;
        if (p.boards_traj) p.boards_traj[off] = bd;
        if (POLICY == 2 || p.legal_mask) {
          const uint32_t lm = legal_mask(bd.x, bd.y, bd.z, bd.w);
          if (p.legal_mask) p.legal_mask[off] = (uint8_t)lm;
          if (POLICY == 2) mask = lm;                  // what the next step's draw chooses among
        }
      }
    }
    p.boards[i] = bd;
  }
}

// Game2048Env.reset (:102-111)
__global__ void __launch_bounds__(kThreads)
g2048_reset_kernel(uint4* boards, const uint8_t* reset_mask, uint64_t n, uint64_t env_id_base, uint64_t seed,
                   uint64_t reset_index) {
  __shared__ Board4 s_lut[32];
  const Board4* lut = make_reset_lut(s_lut);
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  DrawStream draws(seed, reset_index, TAG_RESET);
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    if (reset_mask && !reset_mask[i]) continue;
    const Words w = draws.words(env_id_base + i);
    uint4 bd;
    fresh_board(lut, w.w1, w.w2, bd.x, bd.y, bd.z, bd.w);
    boards[i] = bd;
  }
}

// Game2048Env.add_tile (:166-176)
__global__ void __launch_bounds__(kThreads)
g2048_add_tile_kernel(uint4* boards, uint64_t n, uint64_t env_id_base, uint64_t seed, uint64_t step_index) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  DrawStream draws(seed, step_index, TAG_STEP);
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    uint4 bd = boards[i];
    const Words w = draws.words(env_id_base + i);
    spawn(bd.x, bd.y, bd.z, bd.w, w.w0);
    boards[i] = bd;
  }
}

// Game2048Env.move(direction, trial) (:194-241), no spawn
__global__ void __launch_bounds__(kThreads)
g2048_move_kernel(const uint4* in, uint4* out, const uint8_t* directions, uint32_t* scores, uint8_t* changed,
                  uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    uint4 bd = in[i];
    const uint32_t act = directions[i] & 3u;
    uint32_t a, b, c, d;
    orient(kOrientIn[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
    const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
    const uint32_t s = (uint32_t)slide_merge(a, b, c, d);
    orient(kOrientOut[act], a, b, c, d, bd.x, bd.y, bd.z, bd.w);
    if (out) out[i] = bd;
    if (scores) scores[i] = s;
    if (changed) changed[i] = (((a ^ a0) | (b ^ b0)) | ((c ^ c0) | (d ^ d0))) != 0u ? 1 : 0;
  }
}

// legal mask / highest / empties / isend (:186-192, :262-280)
__global__ void __launch_bounds__(kThreads)
g2048_status_kernel(const uint4* boards, uint8_t* lm, uint8_t* hi, uint8_t* ne, uint8_t* is_end,
                    uint32_t max_tile_exp, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const uint4 bd = boards[i];
    const uint32_t empties = count_empty(bd.x, bd.y, bd.z, bd.w);
    const uint32_t h = highest_exp(bd.x, bd.y, bd.z, bd.w);
    if (lm) lm[i] = (uint8_t)legal_mask(bd.x, bd.y, bd.z, bd.w);
    if (hi) hi[i] = (uint8_t)h;
    if (ne) ne[i] = (uint8_t)empties;
    if (is_end)
      is_end[i] = ((max_tile_exp != 0u && h == max_tile_exp) ||
                   (empties == 0u && full_board_is_dead(bd.x, bd.y, bd.z, bd.w))) ? 1 : 0;
  }
}

// stack() (:17-32): obs[n][16][4][4].  The output is cut into 16-byte chunks and every thread writes exactly
// one of them with one 128-bit store, consecutive threads consecutive chunks (512 contiguous bytes per warp
// instruction) whatever the dtype:
//   u8   16 chunks per board: a whole channel (16 cells)     bf16 32: two rows of a channel (8 cells)
//   f32  64: one row of a channel (4 cells)                  i64 128: half a row (2 cells)
// The board rows a chunk needs are re-read through L1 (16 to 128 threads share a board).
// eq_flags(r, ch): 0/1 per byte, 1 where the cell holds exponent ch (exponents >= 16 match no channel,
// :28-30; ch 0 = empty, :25).
__device__ __forceinline__ uint32_t eq_flags(uint32_t r, uint32_t ch) { return (~((r ^ (ch * K1)) + L7) & H) >> 7; }
struct bf16x4 { uint32_t lo, hi; };   // tag type of the bf16 encoder
template <typename T> struct ObsChunk;
template <> struct ObsChunk<uint8_t> {
  static constexpr uint32_t kPerBoardLog2 = 4;
  static __device__ __forceinline__ uint4 make(const uint32_t* board, uint32_t c) {
    const uint4 b = *reinterpret_cast<const uint4*>(board);
    return make_uint4(eq_flags(b.x, c), eq_flags(b.y, c), eq_flags(b.z, c), eq_flags(b.w, c));
  }
};
template <> struct ObsChunk<bf16x4> {
  static constexpr uint32_t kPerBoardLog2 = 5;
  static __device__ __forceinline__ uint32_t pair(uint32_t f) {       // two 0/1 bytes -> two bf16 (1.0 = 0x3F80)
    return ((f & 1u) * 0x3F80u) | (((f >> 8) & 1u) * 0x3F800000u);
  }
  static __device__ __forceinline__ uint4 make(const uint32_t* board, uint32_t c) {
    const uint2 r = *reinterpret_cast<const uint2*>(board + 2u * (c & 1u));
    const uint32_t f0 = eq_flags(r.x, c >> 1), f1 = eq_flags(r.y, c >> 1);
    return make_uint4(pair(f0), pair(f0 >> 16), pair(f1), pair(f1 >> 16));
  }
};
template <> struct ObsChunk<float> {
  static constexpr uint32_t kPerBoardLog2 = 6;
  static __device__ __forceinline__ uint4 make(const uint32_t* board, uint32_t c) {
    const uint32_t f = eq_flags(board[c & 3u], c >> 2);
    constexpr uint32_t one = 0x3F800000u;
    return make_uint4((f & 1u) * one, ((f >> 8) & 1u) * one, ((f >> 16) & 1u) * one, (f >> 24) * one);
  }
};
template <> struct ObsChunk<int64_t> {
  static constexpr uint32_t kPerBoardLog2 = 7;
  static __device__ __forceinline__ uint4 make(const uint32_t* board, uint32_t c) {
    const uint32_t f = eq_flags(board[(c >> 1) & 3u], c >> 3) >> (16u * (c & 1u));
    return make_uint4(f & 1u, 0u, (f >> 8) & 1u, 0u);
  }
};

template <typename T>
__global__ void __launch_bounds__(kThreads) g2048_obs_kernel(const uint32_t* boards, uint4* obs, uint64_t n_chunks) {
  constexpr uint32_t kLog2 = ObsChunk<T>::kPerBoardLog2;
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t t = (uint64_t)blockIdx.x * kThreads + threadIdx.x; t < n_chunks; t += stride)
    obs[t] = ObsChunk<T>::make(boards + (t >> kLog2) * 4u, (uint32_t)t & ((1u << kLog2) - 1u));
}

__global__ void __launch_bounds__(kThreads)
g2048_values_from_exp_kernel(const uint8_t* exps, int64_t* values, uint64_t n_cells) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n_cells; i += stride) {
    const uint32_t e = exps[i];
    values[i] = e ? (int64_t)(1ull << (e & 63u)) : 0;
  }
}

// The same four cells (one board row) per thread: one 32-bit load, two 128-bit stores.
__global__ void __launch_bounds__(kThreads)
g2048_values_from_exp4_kernel(const uint32_t* rows, longlong2* values, uint64_t n_rows) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n_rows; i += stride) {
    const uint32_t r = rows[i];
    auto val = [](uint32_t e) { return e ? (long long)(1ull << (e & 63u)) : 0ll; };
    values[2 * i] = make_longlong2(val(r & 0xFFu), val((r >> 8) & 0xFFu));
    values[2 * i + 1] = make_longlong2(val((r >> 16) & 0xFFu), val(r >> 24));
  }
}

__device__ __forceinline__ uint32_t exp_of_value(int64_t v, uint32_t* bad_count) {
  uint32_t e = 0;
  if (v != 0) {
    const bool ok = v >= 2 && v <= (1ll << 31) && (v & (v - 1)) == 0;
    if (ok) e = 63u - (uint32_t)__clzll(v);
    else if (bad_count) atomicAdd(bad_count, 1u);
  }
  return e;
}
__global__ void __launch_bounds__(kThreads)
g2048_exp_from_values4_kernel(const longlong2* values, uint32_t* rows, uint64_t n_rows, uint32_t* bad_count) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n_rows; i += stride) {
    const longlong2 a = values[2 * i], b = values[2 * i + 1];
    rows[i] = exp_of_value(a.x, bad_count) | (exp_of_value(a.y, bad_count) << 8) |
              (exp_of_value(b.x, bad_count) << 16) | (exp_of_value(b.y, bad_count) << 24);
  }
}

__global__ void __launch_bounds__(kThreads)
g2048_exp_from_values_kernel(const int64_t* values, uint8_t* exps, uint64_t n_cells, uint32_t* bad_count) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n_cells; i += stride) {
    exps[i] = (uint8_t)exp_of_value(values[i], bad_count);
  }
}

__global__ void __launch_bounds__(kThreads)
g2048_philox_kernel(const uint4* ctr, uint32_t k0, uint32_t k1, uint4* out, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const uint4 c = ctr[i];
    const Words w = philox4x32_10(c.x, c.y, c.z, c.w, k0, k1);
    out[i] = make_uint4(w.w0, w.w1, w.w2, w.w3);
  }
}

__global__ void __launch_bounds__(kThreads)
g2048_philox2x32_kernel(const uint2* ctr, uint32_t key, uint2* out, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const uint2 c = ctr[i];
    const Pair x = philox2x32_10(c.x, c.y, key);
    out[i] = make_uint2(x.x0, x.x1);
  }
}

__global__ void __launch_bounds__(kThreads)
g2048_draw_words_kernel(uint4* out, uint64_t n, uint64_t env_id_base, uint64_t seed, uint64_t idx, uint32_t tag) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  DrawStream draws(seed, idx, tag);
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const Words w = draws.words(env_id_base + i);
    out[i] = make_uint4(w.w0, w.w1, w.w2, w.w3);
  }
}

// ------------------------------------------------------------------------------------
// One env, one launch (g2048_one): the single-env class of the reference (Game2048Env.step/reset/move/add_tile/
// isend/highest/stack on ONE 4x4 Matrix of tile values) as a single kernel over a packed in/out block.  The block
// may live in pinned host memory (the device reads and writes it over PCIe: zero copy), so a call is one launch
// and one stream synchronisation — the per-call latency floor for train.py:150-165 / gather_training_data.py
// :141-145, which step one env at a time.
// ------------------------------------------------------------------------------------
struct OneParams {
  G2048OneIO* io;
  int op;
  uint32_t action;
  uint32_t trial;
  uint64_t seed, index;
  float illegal_move_reward;
  uint32_t max_tile_exp;
};

__global__ void __launch_bounds__(256) g2048_one_kernel(const OneParams p) {
  __shared__ long long s_vals[16];
  __shared__ Board4 s_lut[32];
  const Board4* lut = make_reset_lut(s_lut);
  G2048OneIO* io = p.io;
  if (threadIdx.x == 0) {
    uint32_t r[4] = {0u, 0u, 0u, 0u};
    uint32_t bad = 0u;
    long long vals[16];
    for (int c = 0; c < 16; ++c) {
      const long long v = p.op == G2048_ONE_RESET ? 0ll : (long long)io->values[c];
      vals[c] = v;
      uint32_t e = 0u;
      if (v != 0) {
        if (v >= 2 && v <= (1ll << 31) && (v & (v - 1)) == 0) e = 63u - (uint32_t)__clzll(v);
        else ++bad;
      }
      r[c >> 2] |= e << (8 * (c & 3));
    }
    float reward = 0.f;
    uint32_t score = 0u, done = 0u, illegal = 0u, changed = 0u;
    bool write = false;
    if (bad == 0u || p.op == G2048_ONE_RESET) {
      switch (p.op) {
        case G2048_ONE_STEP: {                                                           // :76-100, no auto-reset
          const Words w = draw_words(p.seed, 0ull, p.index, TAG_STEP);
          const StepOut o = step_board(lut, r[0], r[1], r[2], r[3], p.action & 3u, w, p.max_tile_exp, true, false);
          score = (uint32_t)o.score;
          reward = o.legal ? o.score : p.illegal_move_reward;
          done = o.done ? 1u : 0u;
          illegal = o.legal ? 0u : 1u;
          changed = o.legal ? 1u : 0u;
          write = o.legal;
          break;
        }
        case G2048_ONE_RESET: {                                                          // :102-111
          const Words w = draw_words(p.seed, 0ull, p.index, TAG_RESET);
          fresh_board(lut, w.w1, w.w2, r[0], r[1], r[2], r[3]);
          write = true;
          break;
        }
        case G2048_ONE_MOVE: {                                                           // :194-241
          const uint32_t act = p.action & 3u;
          uint32_t a, b, c, d, m0, m1, m2, m3;
          orient(kOrientIn[act], r[0], r[1], r[2], r[3], a, b, c, d);
          const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
          score = (uint32_t)slide_merge(a, b, c, d);
          reward = (float)score;
          changed = ((((a ^ a0) | (b ^ b0)) | ((c ^ c0) | (d ^ d0))) != 0u) ? 1u : 0u;
          orient(kOrientOut[act], a, b, c, d, m0, m1, m2, m3);
          if (changed && !p.trial) { r[0] = m0; r[1] = m1; r[2] = m2; r[3] = m3; write = true; }
          break;
        }
        case G2048_ONE_ADD_TILE: {                                                       // :166-176
          const Words w = draw_words(p.seed, 0ull, p.index, TAG_STEP);
          changed = spawn(r[0], r[1], r[2], r[3], w.w0) != 0u ? 1u : 0u;
          write = changed != 0u;
          break;
        }
        default: break;                                                                  // G2048_ONE_STATUS: queries only
      }
    }
    if (write) {
      for (int c = 0; c < 16; ++c) {
        const uint32_t e = (r[c >> 2] >> (8 * (c & 3))) & 0xFFu;
        vals[c] = e ? (long long)(1ull << (e & 63u)) : 0ll;
        io->values[c] = vals[c];
      }
    }
    for (int c = 0; c < 16; ++c) s_vals[c] = vals[c];
    const uint32_t h = highest_exp(r[0], r[1], r[2], r[3]);
    const uint32_t empties = count_empty(r[0], r[1], r[2], r[3]);
    io->reward = reward;
    io->score = score;
    io->done = (uint8_t)done;
    io->illegal = (uint8_t)illegal;
    io->changed = (uint8_t)changed;
    io->highest_exp = (uint8_t)h;                                                        // :190-192
    io->legal_mask = (uint8_t)legal_mask(r[0], r[1], r[2], r[3]);
    io->n_empty = (uint8_t)empties;                                                      // :186-188
    io->is_end = ((p.max_tile_exp != 0u && h == p.max_tile_exp) ||                         // :262-280
                  (empties == 0u && full_board_is_dead(r[0], r[1], r[2], r[3]))) ? 1 : 0;
    io->bad_cells = (uint8_t)(bad > 255u ? 255u : bad);
  }
  __syncthreads();
  // stack() (:17-32) of the Matrix as it is now, straight from the tile VALUES: channel 0 = empty, channel k =
  // (cell == 2^k); a cell holding anything else lights no channel, exactly like the reference's comparison.
  const uint32_t ch = threadIdx.x >> 4, cell = threadIdx.x & 15u;
  const long long v = s_vals[cell];
  io->obs[threadIdx.x] = (ch == 0u) ? (v == 0 ? 1 : 0) : (v == (1ll << ch) ? 1 : 0);
}

int launch_check(const char* name) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, name);
  return G2048_OK;
}

// ------------------------------------------------------------------------------------
// stateful host-buffer environment
// ------------------------------------------------------------------------------------
}  // namespace g2048

// ------------------------------------------------------------------------------------
// G2048_BOARDS_BYTES_PACKED_WIRE: the host half.  The boards of a step arrive in pinned staging memory 4 bits per
// cell, slice by slice; a small pool of host threads expands every slice into the caller's [n*16] byte array while
// the next slice is still on the wire.  The pool sleeps between steps (condition variable) and spins within one.
// ------------------------------------------------------------------------------------
// 8 bytes -> 16 bytes per board: byte b of the packed board holds cells 2b (low nibble) and 2b+1 (high nibble).
#ifndef G2048_UNPACK_NT       // 1: non-temporal stores into the caller's array.  Measured (profiles/r02_e2e_wire.log): SLOWER,
#define G2048_UNPACK_NT 0    // 0.46 vs 0.34 ms per 1 Mi boards — with ordinary stores the caller's 16 MB array stays in the
#endif                       // host's last-level cache from step to step and the expansion never waits for DRAM
static void unpack_nibble_boards(const uint8_t* src, uint8_t* dst, size_t boards) {
  size_t i = 0;
#if defined(__SSE2__)
  const __m128i low = _mm_set1_epi8(0x0F);
  const bool aligned = G2048_UNPACK_NT && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
  for (; i + 2 <= boards; i += 2) {
    const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 8 * i));       // two boards
    const __m128i lo = _mm_and_si128(x, low), hi = _mm_and_si128(_mm_srli_epi16(x, 4), low);
    const __m128i b0 = _mm_unpacklo_epi8(lo, hi), b1 = _mm_unpackhi_epi8(lo, hi);
    __m128i* out = reinterpret_cast<__m128i*>(dst + 16 * i);
    if (aligned) { _mm_stream_si128(out, b0); _mm_stream_si128(out + 1, b1); }
    else { _mm_storeu_si128(out, b0); _mm_storeu_si128(out + 1, b1); }
  }
  if (aligned) _mm_sfence();
#endif
  for (; i < boards; ++i)
    for (int b = 0; b < 8; ++b) {
      const uint8_t v = src[8 * i + b];
      dst[16 * i + 2 * b] = v & 15u;
      dst[16 * i + 2 * b + 1] = v >> 4;
    }
}

struct UnpackPool {
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv;
  uint64_t epoch = 0;                 // guarded by m: a new job is published by incrementing it
  bool quit = false;                  // guarded by m
  // the job of the current epoch (written by the stepping thread before the epoch is published)
  const uint8_t* src = nullptr;
  uint8_t* dst = nullptr;
  uint64_t lo[64], cnt[64];
  int n_slices = 0;
  std::atomic<int> ready[64];         // slice c: 0 = still on the wire, 1 = in `src`, -1 = the step failed, give up
  std::atomic<int> done{0};           // workers that are through with the current job

  void run(int t, int T) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> hold(m);
        cv.wait(hold, [&] { return quit || epoch != seen; });
        if (quit) return;
        seen = epoch;
      }
      for (int c = 0; c < n_slices; ++c) {
        int r;
        while ((r = ready[c].load(std::memory_order_acquire)) == 0) {
#if defined(__SSE2__)
          _mm_pause();
#endif
        }
        if (r < 0) break;
        // this worker's share of the slice, in pairs of boards
        const uint64_t pairs = (cnt[c] + 1) / 2, a = pairs * t / T * 2, b = pairs * (t + 1) / T * 2;
        const uint64_t hi = b < cnt[c] ? b : cnt[c];
        if (a < hi) unpack_nibble_boards(src + 8 * (lo[c] + a), dst + 16 * (lo[c] + a), hi - a);
      }
      done.fetch_add(1, std::memory_order_release);
    }
  }
  explicit UnpackPool(int T) {
    for (int c = 0; c < 64; ++c) ready[c].store(0);
    for (int t = 0; t < T; ++t) workers.emplace_back([this, t, T] { run(t, T); });
  }
  ~UnpackPool() {
    { std::lock_guard<std::mutex> hold(m); quit = true; }
    cv.notify_all();
    for (std::thread& w : workers) w.join();
  }
  // Publish a job: every slice still on the wire.
  void start(const uint8_t* s_, uint8_t* d_, int slices) {
    src = s_; dst = d_; n_slices = slices;
    for (int c = 0; c < slices; ++c) ready[c].store(0, std::memory_order_relaxed);
    done.store(0, std::memory_order_relaxed);
    { std::lock_guard<std::mutex> hold(m); ++epoch; }
    cv.notify_all();
  }
  void finish() {                      // wait until every worker is through with the job
    while (done.load(std::memory_order_acquire) != (int)workers.size()) std::this_thread::yield();
  }
  void abort_from(int c) {             // the step failed: release the workers from the slices that will never arrive
    for (; c < n_slices; ++c) ready[c].store(-1, std::memory_order_release);
  }
};

static int default_unpack_threads() {
  int cpus = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = CPU_COUNT(&set);      // a rank pinned to 4 cores gets 2 threads, not 8
#endif
  // half the CPUs on a big host (the expansion is bound by the caches, not by instructions: eight threads are
  // enough); all but one — the stepping thread spins on the slice events — on a small one
  int t = cpus > 8 ? cpus / 2 : cpus - 1;
  return t < 1 ? 1 : (t > 8 ? 8 : t);
}

struct G2048Env {
  G2048EnvConfig cfg;
  uint64_t step_index, reset_index;
  uint32_t n_chunks;
  // device state / results
  uint8_t* d_boards;
  uint8_t* d_actions;
  float* d_rewards;
  uint8_t* d_dones;
  uint8_t* d_illegal;
  uint8_t* d_highest;
  uint8_t* d_mask;
  uint32_t* d_ep_score;
  uint32_t* d_ep_len;
  uint8_t* d_nibble;            // [n*8] compact host format (cfg.board_format == G2048_BOARDS_NIBBLE)
  uint32_t* d_overflow;         // [64] per-slice counters of boards that do not fit the compact format
  uint32_t* h_overflow;         // pinned mirror of d_overflow
  uint8_t* h_nibble;            // [n*8] pinned staging of the packed boards (cfg.board_format == G2048_BOARDS_BYTES_PACKED_WIRE)
  UnpackPool* pool;             // the threads that expand them into the caller's array
  cudaStream_t streams[4];
  int n_streams;
  cudaEvent_t slice_done[64];   // slice c's kernel has finished (the small result copies wait for it on another stream)
  int n_events;
};

using namespace g2048;

extern "C" {

int g2048_abi_version(void) { return G2048_ABI_VERSION; }
const char* g2048_last_error(void) { return g_err; }

// The boards [lo, lo+m) of a call as a call of their own (every per-board pointer advanced).
static G2048StepArgs slice_args(const G2048StepArgs& a, uint64_t lo, uint64_t m) {
  G2048StepArgs s = a;
  auto adv = [lo](auto*& ptr, uint64_t elems_per_board) { if (ptr) ptr += lo * elems_per_board; };
  adv(s.boards, 16); adv(s.boards_out, 16); adv(s.actions, 1); adv(s.rewards, 1); adv(s.dones, 1);
  adv(s.illegal, 1); adv(s.highest_exp, 1); adv(s.legal_mask, 1); adv(s.terminal_boards, 16);
  adv(s.ep_score, 1); adv(s.ep_len, 1); adv(s.final_score, 1); adv(s.final_len, 1); adv(s.forced_draws, 4);
  adv(s.ep_return, 1); adv(s.final_return, 1); adv(s.boards_nibble, 8);
  s.n = m;
  s.env_id_base = a.env_id_base + lo;
  return s;
}

// The step kernels the library instantiates: (output set, device-side step counter) -> kernel.  Any other
// combination of optional pointers runs the generic kernel.
struct StepKernelEntry { uint32_t out; bool counter; int policy; const void* fn; };
static const StepKernelEntry kStepKernels[] = {
    {0u, false, 0, (const void*)g2048_step_kernel<0u, false>},
    {0u, true, 0, (const void*)g2048_step_kernel<0u, true>},
    {O_MASK, false, 0, (const void*)g2048_step_kernel<O_MASK, false>},
    {O_MASK, true, 0, (const void*)g2048_step_kernel<O_MASK, true>},
    {O_EPRUN, false, 0, (const void*)g2048_step_kernel<O_EPRUN, false>},
    {O_EPRUN | O_NIBBLE, false, 0, (const void*)g2048_step_kernel<O_EPRUN | O_NIBBLE, false>},
    {kOutEval, false, 0, (const void*)g2048_step_kernel<kOutEval, false>},
    {kOutAll, false, 0, (const void*)g2048_step_kernel<kOutAll, false>},
    {O_GENERIC, false, 0, (const void*)g2048_step_kernel<O_GENERIC, false>},
    {O_GENERIC, true, 0, (const void*)g2048_step_kernel<O_GENERIC, true>},
#if !G2048_TMA && !G2048_PIPELINE
    // the kernel draws the actions itself (G2048_FLAG_POLICY_UNIFORM / _LEGAL)
    {0u, false, 1, (const void*)g2048_step_kernel<0u, false, 1>},
    {O_MASK, false, 2, (const void*)g2048_step_kernel<O_MASK, false, 2>},
    {O_MASK, true, 2, (const void*)g2048_step_kernel<O_MASK, true, 2>},
    {O_GENERIC, false, 1, (const void*)g2048_step_kernel<O_GENERIC, false, 1>},
    {O_GENERIC, false, 2, (const void*)g2048_step_kernel<O_GENERIC, false, 2>},
    {O_GENERIC, true, 1, (const void*)g2048_step_kernel<O_GENERIC, true, 1>},
    {O_GENERIC, true, 2, (const void*)g2048_step_kernel<O_GENERIC, true, 2>},
#endif
};
// The optional pointers of a call as an output set; *exact = false when a pointer group is only partly given
// (the specialised kernels write whole groups).
static uint32_t output_set(const G2048StepArgs* a, bool* exact) {
  uint32_t have = 0u;
  *exact = true;
  if (a->illegal) have |= O_ILLEGAL;
  if (a->highest_exp) have |= O_HIGHEST;
  if (a->legal_mask) have |= O_MASK;
  if (a->terminal_boards) have |= O_TERMINAL;
  if (a->forced_draws) have |= O_FORCED;
  if (a->boards_nibble) have |= O_NIBBLE;
  auto group = [&](const void* x, const void* y, uint32_t bit) {
    if (x && y) have |= bit;
    else if (x || y) { have |= bit; *exact = false; }
  };
  group(a->ep_score, a->ep_len, O_EPRUN);
  group(a->final_score, a->final_len, O_EPFINAL);
  group(a->ep_return, a->final_return, O_EPRET);
  return have;
}
static const void* pick_step_kernel(const G2048StepArgs* a) {
  bool exact = true;
  const uint32_t have = output_set(a, &exact);
  const bool counter = a->step_counter != nullptr;
  const int policy = (a->flags & G2048_FLAG_POLICY_LEGAL) ? 2 : (a->flags & G2048_FLAG_POLICY_UNIFORM) ? 1 : 0;
  const void* generic = nullptr;
  for (const StepKernelEntry& e : kStepKernels) {
    if (e.counter != counter || e.policy != policy) continue;
    if (exact && e.out == have) return e.fn;
    if (e.out == O_GENERIC) generic = e.fn;
  }
  return generic;
}

static void fill_step_params(const G2048StepArgs* a, bool bump_counter, StepParams& p) {
  std::memset(&p, 0, sizeof p);
  p.boards = reinterpret_cast<const uint4*>(a->boards);
  p.boards_out = reinterpret_cast<uint4*>(a->boards_out ? a->boards_out : a->boards);
  p.actions = a->actions;
  p.rewards = a->rewards;
  p.dones = a->dones;
  p.illegal = a->illegal;
  p.highest_exp = a->highest_exp;
  p.legal_mask = a->legal_mask;
  p.terminal_boards = reinterpret_cast<uint4*>(a->terminal_boards);
  p.ep_score = a->ep_score;
  p.ep_len = a->ep_len;
  p.final_score = a->final_score;
  p.final_len = a->final_len;
  p.ep_return = a->ep_return;
  p.final_return = a->final_return;
  p.boards_nibble = reinterpret_cast<uint2*>(a->boards_nibble);
  p.nibble_overflow = a->nibble_overflow;
  p.forced_draws = reinterpret_cast<const uint4*>(a->forced_draws);
  p.step_counter = a->step_counter;
  p.chain = reinterpret_cast<unsigned long long*>(a->chain);
  p.chain_tag = a->step_index + 1ull;
  p.n = (uint32_t)a->n;
  p.env_lo = (uint32_t)a->env_id_base;
  p.env_hi = (uint32_t)(a->env_id_base >> 32);
  p.seed = a->seed;
  make_stream_keys(stream_key(a->seed, a->step_index, a->env_id_base, TAG_STEP), (uint32_t)a->step_index, p.keys);
  if (a->flags & (G2048_FLAG_POLICY_UNIFORM | G2048_FLAG_POLICY_LEGAL))
    make_stream_keys(stream_key(a->seed, a->step_index, a->env_id_base, TAG_POLICY), (uint32_t)a->step_index, p.policy_keys);
  p.actions_out = const_cast<uint8_t*>(a->actions);
  p.illegal_move_reward = a->illegal_move_reward;
  p.max_tile_exp = a->max_tile_exp;
  p.flags = (a->flags & ~(kFlagBumpCounter | kFlagPrefetch | kFlagChained | G2048_FLAG_CHAINED | G2048_FLAG_CHAIN_INTERLEAVED)) |
            (bump_counter ? kFlagBumpCounter : 0u) | (a->n >= G2048_PREFETCH_MIN_N ? kFlagPrefetch : 0u) |
            (a->chain && (a->flags & G2048_FLAG_CHAINED) && a->step_index != 0 ? kFlagChained : 0u);
}

// Launch configuration of a step over n boards (shape_for), with the kernel attributes it relies on set once per
// device and host thread.
// chain: 0 = no chain buffer, 1 = chain buffer, 2 = chain buffer + G2048_FLAG_CHAIN_INTERLEAVED
static void step_launch_config(uint64_t n, int chain, cudaStream_t s, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr) {
  std::memset(&cfg, 0, sizeof cfg);
  if (chain) {
    // Launches that carry a chain buffer overlap on an SM: 256-thread CTAs (five fit an SM at 48 registers).
    //   a chain stepped back to back: four CTAs per SM — only ONE launch is at work at a time (its successor's
    //     warps wait), so the launch itself must fill the SM; the fifth place is where a CTA of the next launch
    //     runs its prologue and waits (1 Mi boards 10.5 -> 9.4 us, 262,144 4.1 -> 3.6 us);
    //   interleaved chains (several env sets round-robin): one CTA per SM — consecutive launches are independent,
    //     five of them share an SM side by side and every CTA lives four times as long, which amortises its start
    //     and its end (1 Mi boards 10.8 -> 9.05 us, 262,144 4.4 -> 2.5 us; profiles/r02_chain_shapes.log).
    //     A launch that alone would run more than kChainNarrowMaxIters boards per thread gets two CTAs per SM: it
    //     costs 2 % in steady state (1 Mi boards: 9.23 vs 9.03 us) but its drain at a synchronisation point — an
    //     event every 20 launches in the driver's benchmark — is half as long (9.78 vs 10.50 us per step there;
    //     profiles/r02_chain_bench_shapes.log).  The shape depends on n and the flag only: it is the chain's.
    const uint64_t need = (n + kChainThreads - 1) / kChainThreads;
    uint64_t per_sm = kChainCtasPerSm;
    if (chain == 2)
      per_sm = n > (uint64_t)sm_count() * kChainThreads * kChainNarrowCtasPerSm * kChainNarrowMaxIters ? 2u * kChainNarrowCtasPerSm
                                                                                                    : kChainNarrowCtasPerSm;
    uint64_t cap = (uint64_t)sm_count() * per_sm;
    if (cap > kChainSlices) cap = kChainSlices;
    cfg.gridDim = dim3((unsigned)(need < cap ? need : cap));
    cfg.blockDim = dim3(kChainThreads);
  } else {
#if G2048_TMA || !G2048_PAIR_LUT
  {
    const uint64_t need = (n + kStepThreads - 1) / kStepThreads, cap = (uint64_t)sm_count() * kStepCtasPerSm;
    cfg.gridDim = dim3((unsigned)(need < cap ? need : cap));
  }
  cfg.blockDim = dim3(kStepThreads);
#elif G2048_PERSISTENT
  const LaunchShape shape = shape_for(n, sizeof(PairLut) + 256, true);
  cfg.gridDim = dim3(shape.grid);
  cfg.blockDim = dim3(shape.block);
  cfg.dynamicSmemBytes = shape.pad_smem;
#else
  cfg.gridDim = dim3((unsigned)((n + kStepThreads - 1) / kStepThreads));
  cfg.blockDim = dim3(kStepThreads);
#endif
  }
  {
    static std::atomic<uint64_t> seen{0};
    static std::mutex lock;
    once_per_device(seen, lock, [] {                   // allow the padding, prefer shared memory
      for (const StepKernelEntry& e : kStepKernels) {
#if G2048_TMA
        cudaFuncSetAttribute(e.fn, cudaFuncAttributePreferredSharedMemoryCarveout, 50);   // a ring per CTA, two CTAs per SM
#else
        cudaFuncSetAttribute(e.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxPadSmem);
        cudaFuncSetAttribute(e.fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
#endif
      }
    });
  }
  cfg.stream = s;
#if G2048_PDL
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#else
  (void)attr;
#endif
}

// One launch; the env ids of the call do not cross a multiple of 2^32.
static int launch_step(const G2048StepArgs* a, cudaStream_t s, bool bump_counter) {
  StepParams p;
  fill_step_params(a, bump_counter, p);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  step_launch_config(a->n, a->chain ? ((a->flags & G2048_FLAG_CHAIN_INTERLEAVED) ? 2 : 1) : 0, s, cfg, attr);
  void* kargs[] = {&p};
  const void* kernel = pick_step_kernel(a);
  if (!kernel) return fail(G2048_ERR_INVALID, "g2048_step: this build has no kernel for the requested policy flag");
  const cudaError_t le = cudaLaunchKernelExC(&cfg, kernel, kargs);
  if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(g2048_step_kernel)");
  return G2048_OK;
}

static int check_step_args(const G2048StepArgs* a, const char* fn) {
  if (!a) return fail(G2048_ERR_INVALID, "%s: args is NULL", fn);
  if (a->n == 0) return G2048_OK;
  if (!a->boards || !a->actions || !a->rewards || !a->dones)
    return fail(G2048_ERR_INVALID, "%s: boards, actions, rewards and dones are required", fn);
  if (!aligned16(a->boards) || !aligned16(a->boards_out) || !aligned16(a->terminal_boards) ||
      !aligned16(a->forced_draws))
    return fail(G2048_ERR_ALIGN, "%s: boards / boards_out / terminal_boards / forced_draws must be 16-byte aligned", fn);
  if (reinterpret_cast<uintptr_t>(a->boards_nibble) & 7u)
    return fail(G2048_ERR_ALIGN, "%s: boards_nibble must be 8-byte aligned", fn);
  if (a->max_tile_exp > 63u) return fail(G2048_ERR_INVALID, "%s: max_tile_exp %u > 63", fn, a->max_tile_exp);
  if (a->flags & ~(G2048_FLAG_AUTO_RESET | G2048_FLAG_POLICY_UNIFORM | G2048_FLAG_POLICY_LEGAL | G2048_FLAG_CHAINED |
                   G2048_FLAG_CHAIN_INTERLEAVED))
    return fail(G2048_ERR_INVALID, "%s: unknown flags 0x%x", fn, a->flags);
  if ((a->flags & (G2048_FLAG_CHAINED | G2048_FLAG_CHAIN_INTERLEAVED)) && !a->chain)
    return fail(G2048_ERR_INVALID, "%s: G2048_FLAG_CHAINED / G2048_FLAG_CHAIN_INTERLEAVED need the chain buffer", fn);
  if (a->chain && a->step_counter)
    return fail(G2048_ERR_INVALID, "%s: chain and step_counter cannot be combined", fn);
  if (reinterpret_cast<uintptr_t>(a->chain) & 7u)
    return fail(G2048_ERR_ALIGN, "%s: chain must be 8-byte aligned", fn);
  if (a->chain && (G2048_TMA || G2048_PIPELINE))
    return fail(G2048_ERR_INVALID, "%s: this build variant has no chained launches", fn);
  if ((a->flags & G2048_FLAG_POLICY_UNIFORM) && (a->flags & G2048_FLAG_POLICY_LEGAL))
    return fail(G2048_ERR_INVALID, "%s: choose one of G2048_FLAG_POLICY_UNIFORM / G2048_FLAG_POLICY_LEGAL", fn);
  if ((a->flags & G2048_FLAG_POLICY_LEGAL) && !a->legal_mask)
    return fail(G2048_ERR_INVALID, "%s: G2048_FLAG_POLICY_LEGAL draws among the moves in legal_mask, which is NULL", fn);
  if ((a->flags & (G2048_FLAG_POLICY_UNIFORM | G2048_FLAG_POLICY_LEGAL)) && a->forced_draws)
    return fail(G2048_ERR_INVALID, "%s: forced_draws and a policy flag cannot be combined", fn);
  if (a->n > 0xFFFFFF00ull) return fail(G2048_ERR_INVALID, "%s: n must be < 2^32 - 256 per call", fn);
  return G2048_OK;
}

// One validated step.  The kernel treats the high half of the env id as launch-uniform (Philox head): a call whose
// ids cross a multiple of 2^32 is issued as two launches.
static int issue_step(const G2048StepArgs* a, cudaStream_t s) {
  const uint64_t to_boundary = 0x100000000ull - (a->env_id_base & 0xFFFFFFFFull);
  if (a->n > to_boundary) {
    const G2048StepArgs first = slice_args(*a, 0, to_boundary);
    G2048StepArgs second = slice_args(*a, to_boundary, a->n - to_boundary);
    if (second.chain) second.chain += kChainLaunchWords;      // each half chains to the same half of the previous call
    const int rc = launch_step(&first, s, false);
    return rc == G2048_OK ? launch_step(&second, s, true) : rc;     // both read the index, the second advances it
  }
  return launch_step(a, s, true);
}

int g2048_step(const G2048StepArgs* a, void* stream) {
  const int bad = check_step_args(a, "g2048_step");
  if (bad || a->n == 0) return bad;
  const int rc = issue_step(a, static_cast<cudaStream_t>(stream));
  if (rc != G2048_OK) return rc;
  return launch_check("g2048_step_kernel");
}

// n_steps steps, one kernel launch per step, issued back to back from C: the per-launch host cost is one
// cudaLaunchKernelEx (~2 us) instead of a trip through the caller's interpreter (4-6 us per step from Python), which
// is what bounds a batch that a GPU steps in ~3 us (BASELINE config 3 sharded over 8 GPUs: 131,072 boards each).
int g2048_step_n(const G2048StepArgs* a, uint32_t n_steps, uint64_t row_stride, void* stream) {
  const int bad = check_step_args(a, "g2048_step_n");
  if (bad || a->n == 0 || n_steps == 0) return bad;
  if (a->step_counter) return fail(G2048_ERR_INVALID, "g2048_step_n: step_counter must be NULL (the step index is advanced on the host)");
  if (a->boards_out) return fail(G2048_ERR_INVALID, "g2048_step_n: boards_out must be NULL (the steps run in place)");
  if (row_stride != 0 && row_stride < a->n) return fail(G2048_ERR_INVALID, "g2048_step_n: row_stride must be 0 or >= n");
  if (!aligned16(a->terminal_boards ? a->terminal_boards + 16 * row_stride : nullptr))
    return fail(G2048_ERR_ALIGN, "g2048_step_n: terminal_boards rows must stay 16-byte aligned");
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  const DeviceHint pin_device;
  G2048StepArgs k = *a;
  for (uint32_t t = 0; t < n_steps; ++t) {
    const int rc = issue_step(&k, s);
    if (rc != G2048_OK) return rc;
    // per-step arrays move on by one row; per-env state (boards, running episode statistics) stays
    k.actions += row_stride; k.rewards += row_stride; k.dones += row_stride;
    if (k.illegal) k.illegal += row_stride;
    if (k.highest_exp) k.highest_exp += row_stride;
    if (k.legal_mask) k.legal_mask += row_stride;
    if (k.terminal_boards) k.terminal_boards += 16 * row_stride;
    if (k.final_score) k.final_score += row_stride;
    if (k.final_len) k.final_len += row_stride;
    if (k.final_return) k.final_return += row_stride;
    if (k.forced_draws) k.forced_draws += 4 * row_stride;
    if (k.boards_nibble) k.boards_nibble += 8 * row_stride;
    k.step_index += 1;
    if (k.chain) k.flags |= G2048_FLAG_CHAINED;       // step t+1 reads what step t wrote and rows that were there before the call
  }
  return launch_check("g2048_step_kernel");
}

// A caller-built list of steps, issued by one call: element j is a complete g2048_step call (its own boards, rows,
// step index ...), launched in order on `stream`.  The elements may belong to different env sets.  With `events`
// (cudaEvent_t handles): events[0] is recorded before the first launch and events[r] after launch r * every — a whole
// timed run (R regions of `every` launches) is then ONE trip through the caller's interpreter.
static int step_list(const G2048StepArgs* list, uint64_t count, uint64_t every, void* const* events, uint64_t n_events,
                     void* stream, const char* fn) {
  if (count == 0) return G2048_OK;
  if (!list) return fail(G2048_ERR_INVALID, "%s: list is NULL", fn);
  if (n_events && (!events || every == 0)) return fail(G2048_ERR_INVALID, "%s: events need an array and every > 0", fn);
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  const DeviceHint pin_device;
  uint64_t next_event = 0;
  if (n_events) G2048_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(events[next_event++]), s));
  for (uint64_t j = 0; j < count; ++j) {
    const G2048StepArgs* a = list + j;
    const int bad = check_step_args(a, fn);
    if (bad) {
      char first[sizeof g_err];
      std::snprintf(first, sizeof first, "%s", g_err);
      return fail(bad, "%s: element %llu: %s", fn, (unsigned long long)j, first);
    }
    if (a->n != 0) {
      const int rc = issue_step(a, s);
      if (rc != G2048_OK) return rc;
    }
    if (next_event < n_events && (j + 1) % every == 0)
      G2048_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(events[next_event++]), s));
  }
  return launch_check("g2048_step_kernel");
}

int g2048_step_list(const G2048StepArgs* list, uint64_t count, void* stream) {
  return step_list(list, count, 0, nullptr, 0, stream, "g2048_step_list");
}

int g2048_step_list_timed(const G2048StepArgs* list, uint64_t count, uint64_t every, void* const* events,
                          uint64_t n_events, void* stream) {
  return step_list(list, count, every, events, n_events, stream, "g2048_step_list_timed");
}

// lean, with optional outputs, uniform policy, random-legal policy
static const void* const kManyKernels[4] = {
    (const void*)g2048_step_many_kernel<false, 0>, (const void*)g2048_step_many_kernel<true, 0>,
    (const void*)g2048_step_many_kernel<true, 1>, (const void*)g2048_step_many_kernel<true, 2>,
};

// One g2048_step_many launch: no 2^32 boundary of env ids or step indices inside.
static int launch_step_many(const G2048StepManyArgs* a, uint64_t lo, uint64_t m, uint64_t k0, uint64_t ks,
                            cudaStream_t s) {
  ManyParams p;
  const uint64_t row0 = k0 * a->n + lo;
  p.boards = reinterpret_cast<uint4*>(a->boards) + lo;
  p.actions = a->actions ? a->actions + row0 : nullptr;
  p.rewards = a->rewards + row0;
  p.dones = a->dones + row0;
  p.illegal = a->illegal ? a->illegal + row0 : nullptr;
  p.boards_traj = a->boards_traj ? reinterpret_cast<uint4*>(a->boards_traj) + row0 : nullptr;
  p.actions_out = a->actions_out ? a->actions_out + row0 : nullptr;
  p.legal_mask = a->legal_mask ? a->legal_mask + row0 : nullptr;
  p.n = (uint32_t)m;
  p.n_steps = (uint32_t)ks;
  p.row_stride = a->n;
  const uint64_t env0 = a->env_id_base + lo, idx0 = a->step_index + k0;
  p.env_lo = (uint32_t)env0;
  p.idx_lo = (uint32_t)idx0;
  p.key = stream_key(a->seed, idx0, env0, TAG_STEP);
  make_stream_keys(p.key, 0u, p.keys);
  p.policy_key = stream_key(a->seed, idx0, env0, TAG_POLICY);
  make_stream_keys(p.policy_key, 0u, p.policy_keys);
  p.illegal_move_reward = a->illegal_move_reward;
  p.max_tile_exp = a->max_tile_exp;
  p.flags = a->flags;
  const LaunchShape shape = shape_for(m, sizeof(PairLut) + 256, false);
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(shape.grid);
  cfg.blockDim = dim3(shape.block);
  cfg.dynamicSmemBytes = shape.pad_smem;
  cfg.stream = s;
#if G2048_PDL
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#endif
  {
    // once per device: allow the padding of shape_for, prefer shared memory over L1
    static std::atomic<uint64_t> seen{0};
    static std::mutex lock;
    once_per_device(seen, lock, [] {
      for (const void* k : kManyKernels) {
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxPadSmem);
      }
    });
  }
  const bool extras = a->illegal || a->boards_traj || a->legal_mask;
  const int policy = (a->flags & G2048_FLAG_POLICY_LEGAL) ? 2 : (a->flags & G2048_FLAG_POLICY_UNIFORM) ? 1 : 0;
  void* kargs[] = {&p};
  const cudaError_t le = cudaLaunchKernelExC(&cfg, kManyKernels[policy ? 1 + policy : (extras ? 1 : 0)], kargs);
  if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(g2048_step_many_kernel)");
  return G2048_OK;
}

int g2048_step_many(const G2048StepManyArgs* a, void* stream) {
  if (!a) return fail(G2048_ERR_INVALID, "g2048_step_many: args is NULL");
  if (a->n == 0 || a->n_steps == 0) return G2048_OK;
  const bool has_policy = (a->flags & (G2048_FLAG_POLICY_UNIFORM | G2048_FLAG_POLICY_LEGAL)) != 0u;
  if ((a->flags & G2048_FLAG_POLICY_UNIFORM) && (a->flags & G2048_FLAG_POLICY_LEGAL))
    return fail(G2048_ERR_INVALID, "g2048_step_many: choose one of G2048_FLAG_POLICY_UNIFORM / G2048_FLAG_POLICY_LEGAL");
  if (!a->boards || !a->rewards || !a->dones || (!a->actions && !has_policy))
    return fail(G2048_ERR_INVALID, "g2048_step_many: boards, rewards, dones and (without a policy flag) actions are required");
  if (a->actions_out && !has_policy)
    return fail(G2048_ERR_INVALID, "g2048_step_many: actions_out is only written under a policy flag");
  if (!aligned16(a->boards) || !aligned16(a->boards_traj))
    return fail(G2048_ERR_ALIGN, "g2048_step_many: boards / boards_traj must be 16-byte aligned");
  if (a->max_tile_exp > 63u) return fail(G2048_ERR_INVALID, "g2048_step_many: max_tile_exp %u > 63", a->max_tile_exp);
  if (a->flags & ~(G2048_FLAG_AUTO_RESET | G2048_FLAG_POLICY_UNIFORM | G2048_FLAG_POLICY_LEGAL))
    return fail(G2048_ERR_INVALID, "g2048_step_many: unknown flags 0x%x", a->flags);
  if (a->n > 0xFFFFFF00ull) return fail(G2048_ERR_INVALID, "g2048_step_many: n must be < 2^32 - 256 per call");
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  // The kernel treats the high halves of the env id and of the step index as launch-uniform (they are folded
  // into the generator's key): a call that crosses a multiple of 2^32 in either is issued in pieces.  Steps
  // first (a later step reads the boards the earlier one left), boards second.
  for (uint64_t k0 = 0; k0 < a->n_steps;) {
    const uint64_t to_idx = 0x100000000ull - ((a->step_index + k0) & 0xFFFFFFFFull);
    const uint64_t ks = (a->n_steps - k0 < to_idx) ? a->n_steps - k0 : to_idx;
    for (uint64_t lo = 0; lo < a->n;) {
      const uint64_t to_env = 0x100000000ull - ((a->env_id_base + lo) & 0xFFFFFFFFull);
      const uint64_t m = (a->n - lo < to_env) ? a->n - lo : to_env;
      const int rc = launch_step_many(a, lo, m, k0, ks, s);
      if (rc != G2048_OK) return rc;
      lo += m;
    }
    k0 += ks;
  }
  return launch_check("g2048_step_many_kernel");
}

int g2048_reset(uint8_t* boards, const uint8_t* reset_mask, uint64_t n, uint64_t env_id_base, uint64_t seed,
                uint64_t reset_index, void* stream) {
  if (n == 0) return G2048_OK;
  if (!boards) return fail(G2048_ERR_INVALID, "g2048_reset: boards is NULL");
  if (!aligned16(boards)) return fail(G2048_ERR_ALIGN, "g2048_reset: boards must be 16-byte aligned");
  g2048_reset_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint4*>(boards), reset_mask, n, env_id_base, seed, reset_index);
  return launch_check("g2048_reset_kernel");
}

int g2048_add_tile(uint8_t* boards, uint64_t n, uint64_t env_id_base, uint64_t seed, uint64_t step_index,
                   void* stream) {
  if (n == 0) return G2048_OK;
  if (!boards) return fail(G2048_ERR_INVALID, "g2048_add_tile: boards is NULL");
  if (!aligned16(boards)) return fail(G2048_ERR_ALIGN, "g2048_add_tile: boards must be 16-byte aligned");
  g2048_add_tile_kernel<<<grid_for_streaming(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint4*>(boards), n, env_id_base, seed, step_index);
  return launch_check("g2048_add_tile_kernel");
}

int g2048_move(const uint8_t* boards_in, uint8_t* boards_out, const uint8_t* directions, uint32_t* scores,
               uint8_t* changed, uint64_t n, void* stream) {
  if (n == 0) return G2048_OK;
  if (!boards_in || !directions) return fail(G2048_ERR_INVALID, "g2048_move: boards_in and directions are required");
  if (!aligned16(boards_in) || !aligned16(boards_out))
    return fail(G2048_ERR_ALIGN, "g2048_move: boards must be 16-byte aligned");
  g2048_move_kernel<<<grid_for_streaming(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(boards_in), reinterpret_cast<uint4*>(boards_out), directions, scores,
      changed, n);
  return launch_check("g2048_move_kernel");
}

int g2048_status(const uint8_t* boards, uint8_t* legal_mask_out, uint8_t* highest_out, uint8_t* n_empty,
                 uint8_t* is_end, uint32_t max_tile_exp, uint64_t n, void* stream) {
  if (n == 0) return G2048_OK;
  if (!boards) return fail(G2048_ERR_INVALID, "g2048_status: boards is NULL");
  if (!aligned16(boards)) return fail(G2048_ERR_ALIGN, "g2048_status: boards must be 16-byte aligned");
  g2048_status_kernel<<<grid_for_streaming(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(boards), legal_mask_out, highest_out, n_empty, is_end, max_tile_exp, n);
  return launch_check("g2048_status_kernel");
}

int g2048_encode_obs(const uint8_t* boards, void* obs, int dtype, uint64_t n, void* stream) {
  if (n == 0) return G2048_OK;
  if (!boards || !obs) return fail(G2048_ERR_INVALID, "g2048_encode_obs: boards and obs are required");
  if (!aligned16(boards) || !aligned16(obs))
    return fail(G2048_ERR_ALIGN, "g2048_encode_obs: boards and obs must be 16-byte aligned");
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint32_t* b = reinterpret_cast<const uint32_t*>(boards);
  uint4* o = static_cast<uint4*>(obs);
  switch (dtype) {
    case G2048_OBS_U8:   g2048_obs_kernel<uint8_t><<<grid_for_streaming(n << 4), kThreads, 0, s>>>(b, o, n << 4); break;
    case G2048_OBS_BF16: g2048_obs_kernel<bf16x4><<<grid_for_streaming(n << 5), kThreads, 0, s>>>(b, o, n << 5); break;
    case G2048_OBS_F32:  g2048_obs_kernel<float><<<grid_for_streaming(n << 6), kThreads, 0, s>>>(b, o, n << 6); break;
    case G2048_OBS_I64:  g2048_obs_kernel<int64_t><<<grid_for_streaming(n << 7), kThreads, 0, s>>>(b, o, n << 7); break;
    default: return fail(G2048_ERR_INVALID, "g2048_encode_obs: unknown dtype %d", dtype);
  }
  return launch_check("g2048_obs_kernel");
}

int g2048_values_from_exp(const uint8_t* boards, int64_t* values, uint64_t n_cells, void* stream) {
  if (n_cells == 0) return G2048_OK;
  if (!boards || !values) return fail(G2048_ERR_INVALID, "g2048_values_from_exp: NULL pointer");
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((n_cells & 3u) == 0 && (reinterpret_cast<uintptr_t>(boards) & 3u) == 0 && aligned16(values))   // whole rows
    g2048_values_from_exp4_kernel<<<grid_for(n_cells / 4), kThreads, 0, s>>>(
        reinterpret_cast<const uint32_t*>(boards), reinterpret_cast<longlong2*>(values), n_cells / 4);
  else
    g2048_values_from_exp_kernel<<<grid_for(n_cells), kThreads, 0, s>>>(boards, values, n_cells);
  return launch_check("g2048_values_from_exp_kernel");
}

int g2048_exp_from_values(const int64_t* values, uint8_t* boards, uint64_t n_cells, uint32_t* bad_count,
                          void* stream) {
  if (n_cells == 0) return G2048_OK;
  if (!boards || !values) return fail(G2048_ERR_INVALID, "g2048_exp_from_values: NULL pointer");
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((n_cells & 3u) == 0 && (reinterpret_cast<uintptr_t>(boards) & 3u) == 0 && aligned16(values))   // whole rows
    g2048_exp_from_values4_kernel<<<grid_for_streaming(n_cells / 4), kThreads, 0, s>>>(
        reinterpret_cast<const longlong2*>(values), reinterpret_cast<uint32_t*>(boards), n_cells / 4, bad_count);
  else
    g2048_exp_from_values_kernel<<<grid_for(n_cells), kThreads, 0, s>>>(values, boards, n_cells, bad_count);
  return launch_check("g2048_exp_from_values_kernel");
}

int g2048_philox(const uint32_t* ctr, uint32_t key0, uint32_t key1, uint32_t* out, uint64_t n, void* stream) {
  if (n == 0) return G2048_OK;
  if (!ctr || !out) return fail(G2048_ERR_INVALID, "g2048_philox: NULL pointer");
  if (!aligned16(ctr) || !aligned16(out)) return fail(G2048_ERR_ALIGN, "g2048_philox: 16-byte alignment required");
  g2048_philox_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(ctr), key0, key1, reinterpret_cast<uint4*>(out), n);
  return launch_check("g2048_philox_kernel");
}

int g2048_philox2x32(const uint32_t* ctr, uint32_t key, uint32_t* out, uint64_t n, void* stream) {
  if (n == 0) return G2048_OK;
  if (!ctr || !out) return fail(G2048_ERR_INVALID, "g2048_philox2x32: NULL pointer");
  if (((uintptr_t)ctr | (uintptr_t)out) & 7u) return fail(G2048_ERR_ALIGN, "g2048_philox2x32: 8-byte alignment required");
  g2048_philox2x32_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint2*>(ctr), key, reinterpret_cast<uint2*>(out), n);
  return launch_check("g2048_philox2x32_kernel");
}

int g2048_draw_words(uint32_t* words, uint64_t n, uint64_t env_id_base, uint64_t seed, uint64_t index, uint32_t tag,
                     void* stream) {
  if (n == 0) return G2048_OK;
  if (!words) return fail(G2048_ERR_INVALID, "g2048_draw_words: words is NULL");
  if (!aligned16(words)) return fail(G2048_ERR_ALIGN, "g2048_draw_words: words must be 16-byte aligned");
  if (tag > 2u) return fail(G2048_ERR_INVALID, "g2048_draw_words: tag %u is not a stream tag (0 step, 1 reset, 2 policy)", tag);
  g2048_draw_words_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint4*>(words), n, env_id_base, seed, index, tag);
  return launch_check("g2048_draw_words_kernel");
}

int g2048_one(G2048OneIO* io, int op, int action, int trial, uint64_t seed, uint64_t index,
              float illegal_move_reward, uint32_t max_tile_exp, void* stream, int synchronize) {
  if (!io) return fail(G2048_ERR_INVALID, "g2048_one: io is NULL");
  if (reinterpret_cast<uintptr_t>(io) & 7u) return fail(G2048_ERR_ALIGN, "g2048_one: io must be 8-byte aligned");
  if (op < G2048_ONE_STEP || op > G2048_ONE_STATUS) return fail(G2048_ERR_INVALID, "g2048_one: unknown op %d", op);
  if (max_tile_exp > 63u) return fail(G2048_ERR_INVALID, "g2048_one: max_tile_exp %u > 63", max_tile_exp);
  OneParams p;
  p.io = io;
  p.op = op;
  p.action = (uint32_t)action;
  p.trial = trial ? 1u : 0u;
  p.seed = seed;
  p.index = index;
  p.illegal_move_reward = illegal_move_reward;
  p.max_tile_exp = max_tile_exp;
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  g2048_one_kernel<<<1, 256, 0, s>>>(p);
  const int rc = launch_check("g2048_one_kernel");
  if (rc != G2048_OK) return rc;
  if (synchronize) G2048_CUDA(cudaStreamSynchronize(s));
  return G2048_OK;
}

// ---- stateful host-buffer API ---------------------------------------------------------

int g2048_env_destroy(G2048Env* e) {
  if (!e) return G2048_OK;
  cudaSetDevice(e->cfg.device);
  for (int i = 0; i < e->n_streams; ++i) cudaStreamDestroy(e->streams[i]);
  for (int i = 0; i < e->n_events; ++i) cudaEventDestroy(e->slice_done[i]);
  cudaFree(e->d_boards); cudaFree(e->d_actions); cudaFree(e->d_rewards); cudaFree(e->d_dones);
  cudaFree(e->d_illegal); cudaFree(e->d_highest); cudaFree(e->d_mask);
  cudaFree(e->d_ep_score); cudaFree(e->d_ep_len); cudaFree(e->d_nibble); cudaFree(e->d_overflow);
  if (e->h_overflow) cudaFreeHost(e->h_overflow);
  delete e->pool;
  if (e->h_nibble) cudaFreeHost(e->h_nibble);
  delete e;
  return G2048_OK;
}

int g2048_env_create(G2048Env** out, const G2048EnvConfig* cfg) {
  if (!out || !cfg) return fail(G2048_ERR_INVALID, "g2048_env_create: NULL argument");
  if (cfg->n == 0) return fail(G2048_ERR_INVALID, "g2048_env_create: n must be > 0");
  if (cfg->max_tile_exp > 63u) return fail(G2048_ERR_INVALID, "g2048_env_create: max_tile_exp > 63");
  if (cfg->board_format > G2048_BOARDS_BYTES_PACKED_WIRE) return fail(G2048_ERR_INVALID, "g2048_env_create: unknown board_format %u", cfg->board_format);
  if (cfg->unpack_threads > 64u) return fail(G2048_ERR_INVALID, "g2048_env_create: unpack_threads > 64");
  G2048_CUDA(cudaSetDevice(cfg->device));
  G2048Env* e = new (std::nothrow) G2048Env();
  if (!e) return fail(G2048_ERR_NOMEM, "g2048_env_create: out of host memory");
  std::memset(e, 0, sizeof *e);
  e->cfg = *cfg;
  const bool packed_wire = cfg->board_format == G2048_BOARDS_BYTES_PACKED_WIRE;
  // default pipeline depth: a 1/16 lead slice + the rest; with packed boards the rest in three slices, so that the
  // expansion of a slice (host threads) hides behind the transfer of the next
  e->n_chunks = cfg->n_chunks ? cfg->n_chunks : (packed_wire ? 4u : 2u);
  if (e->n_chunks > 64u) e->n_chunks = 64u;
  const uint64_t n = cfg->n;
  cudaError_t err = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(p, bytes); };
  alloc((void**)&e->d_boards, n * 16); alloc((void**)&e->d_actions, n); alloc((void**)&e->d_rewards, n * 4);
  alloc((void**)&e->d_dones, n); alloc((void**)&e->d_illegal, n); alloc((void**)&e->d_highest, n);
  alloc((void**)&e->d_mask, n); alloc((void**)&e->d_ep_score, n * 4); alloc((void**)&e->d_ep_len, n * 4);
  if (cfg->board_format == G2048_BOARDS_NIBBLE || packed_wire) {
    alloc((void**)&e->d_nibble, n * 8); alloc((void**)&e->d_overflow, 64 * 4);
    if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->h_overflow, 64 * 4, cudaHostAllocDefault);
  }
  if (packed_wire) {
    if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->h_nibble, n * 8, cudaHostAllocDefault);
    if (err == cudaSuccess) {
      e->pool = new (std::nothrow) UnpackPool(cfg->unpack_threads ? (int)cfg->unpack_threads : default_unpack_threads());
      if (!e->pool) err = cudaErrorMemoryAllocation;
    }
  }
  if (err == cudaSuccess) err = cudaMemset(e->d_ep_score, 0, n * 4);
  if (err == cudaSuccess) err = cudaMemset(e->d_ep_len, 0, n * 4);
  if (err == cudaSuccess) err = cudaMemset(e->d_boards, 0, n * 16);
  for (int i = 0; i < 4 && err == cudaSuccess; ++i) {
    err = cudaStreamCreateWithFlags(&e->streams[i], cudaStreamNonBlocking);
    if (err == cudaSuccess) e->n_streams = i + 1;
  }
  for (int i = 0; i < 64 && err == cudaSuccess; ++i) {
    err = cudaEventCreateWithFlags(&e->slice_done[i], cudaEventDisableTiming);
    if (err == cudaSuccess) e->n_events = i + 1;
  }
  if (err != cudaSuccess) {
    const int rc = err == cudaErrorMemoryAllocation ? fail(G2048_ERR_NOMEM, "g2048_env_create: %s", cudaGetErrorString(err))
                                                    : cuda_fail(err, "g2048_env_create");
    g2048_env_destroy(e);
    return rc;
  }
  *out = e;
  return G2048_OK;
}

static int env_sync(G2048Env* e) {
  for (int i = 0; i < e->n_streams; ++i) G2048_CUDA(cudaStreamSynchronize(e->streams[i]));
  return G2048_OK;
}

int g2048_env_reset_host(G2048Env* e, uint8_t* boards_host) {
  if (!e) return fail(G2048_ERR_INVALID, "g2048_env_reset_host: env is NULL");
  G2048_CUDA(cudaSetDevice(e->cfg.device));
  const cudaStream_t s = e->streams[0];
  int rc = g2048_reset(e->d_boards, nullptr, e->cfg.n, e->cfg.env_id_base, e->cfg.seed, e->reset_index, s);
  if (rc) return rc;
  e->reset_index += 1;
  G2048_CUDA(cudaMemsetAsync(e->d_ep_score, 0, e->cfg.n * 4, s));
  G2048_CUDA(cudaMemsetAsync(e->d_ep_len, 0, e->cfg.n * 4, s));
  if (boards_host) G2048_CUDA(cudaMemcpyAsync(boards_host, e->d_boards, e->cfg.n * 16, cudaMemcpyDeviceToHost, s));
  G2048_CUDA(cudaStreamSynchronize(s));
  return G2048_OK;
}

int g2048_env_set_boards_host(G2048Env* e, const uint8_t* boards_host) {
  if (!e || !boards_host) return fail(G2048_ERR_INVALID, "g2048_env_set_boards_host: NULL argument");
  G2048_CUDA(cudaSetDevice(e->cfg.device));
  G2048_CUDA(cudaMemcpyAsync(e->d_boards, boards_host, e->cfg.n * 16, cudaMemcpyHostToDevice, e->streams[0]));
  G2048_CUDA(cudaStreamSynchronize(e->streams[0]));
  return G2048_OK;
}

// One step with HOST buffers.  The batch is cut into n_chunks slices of a multiple of 256
// boards (a short lead slice first, see below); slice c runs H2D(actions) -> step kernel -> D2H(results) on stream c % n_streams,
// so the PCIe copies of one slice overlap the kernel and the opposite-direction copies of
// its neighbours.  Pinned caller buffers (cudaHostAlloc / cudaHostRegister) get full DMA
// speed; pageable ones work but are staged by the driver.
int g2048_env_step_host(G2048Env* e, const uint8_t* actions_host, const G2048HostStepOut* o) {
  if (!e || !actions_host || !o) return fail(G2048_ERR_INVALID, "g2048_env_step_host: NULL argument");
  if (!o->boards || !o->rewards || !o->dones)
    return fail(G2048_ERR_INVALID, "g2048_env_step_host: boards, rewards and dones are required");
  G2048_CUDA(cudaSetDevice(e->cfg.device));
  const uint64_t n = e->cfg.n;
  // Chunk schedule: a short lead chunk (n / lead_div boards) so that the first D2H copy starts a few
  // microseconds into the call, then the rest in n_chunks - 1 equal slices whose H2D copies and kernels
  // hide behind the D2H stream of their predecessors.
  // (profiles/r01_e2e_chunks.log: 1/16 lead + 2 slices 0.463 ms per 1 Mi boards, equal thirds 0.488 ms;
  //  profiles/r02_e2e_sweep.log: 1/16 lead + 1 slice 0.451 ms, + 2 slices 0.464 ms: the default is 2 chunks.)
  constexpr uint64_t lead_div = 16;
  const bool packed_wire = e->cfg.board_format == G2048_BOARDS_BYTES_PACKED_WIRE;
  uint64_t lead = 0;
  if (e->n_chunks >= 2 && n >= 65536) lead = (n / lead_div + 255) / 256 * 256;
  // Packed wire, experiment (off): the LAST 1/DIV of the boards travel as plain 16-byte boards straight into the
  // caller's array — they need no expansion, and while they are on the wire the host threads could finish the packed
  // slices.  Measured (profiles/r02_e2e_wire.log): slower for every DIV — 0.347 ms per 1 Mi boards without a plain
  // tail, 0.366 / 0.370 / 0.375 / 0.411 with 1/12, 1/8, 1/6, 1/4 — the extra 8 bytes per board on the wire cost more
  // than the expansion they save.
#ifndef G2048_PACKED_PLAIN_TAIL_DIV
#define G2048_PACKED_PLAIN_TAIL_DIV 0
#endif
  uint64_t tail = 0;
  if (packed_wire && G2048_PACKED_PLAIN_TAIL_DIV && e->n_chunks >= 3 && n >= 65536)
    tail = (n / (G2048_PACKED_PLAIN_TAIL_DIV ? G2048_PACKED_PLAIN_TAIL_DIV : 1) + 255) / 256 * 256;
  const uint64_t n_body = n - tail;                                       // [0, n_body): lead + equal slices; [n_body, n): the tail
  const uint64_t rest_chunks = e->n_chunks - (lead ? 1 : 0) - (tail ? 1 : 0);
  uint64_t per = (n_body - lead + rest_chunks - 1) / rest_chunks;
  per = (per + 255) / 256 * 256;
  const bool packed_format = packed_wire;
  const bool nibble_format = e->cfg.board_format == G2048_BOARDS_NIBBLE || packed_wire;      // what the kernel writes for the wire
  const bool nibble = nibble_format;
#ifndef G2048_E2E_SPLIT_COPIES   // 1: the small result arrays of a slice are copied on a second stream (another copy engine).
#define G2048_E2E_SPLIT_COPIES 0   //    Measured (profiles/r02_e2e_sweep.log): no difference — the call is bound by the board
#endif                             //    bytes over PCIe, not by per-copy set-up — so the simpler schedule ships.
  // One slice: H2D(actions) -> step kernel -> D2H(boards) on stream s; the small result arrays (rewards, dones, ...)
  // follow the kernel on the auxiliary stream, so that their per-copy set-up cost runs beside the board copies
  // instead of between them.
  const cudaStream_t aux = (G2048_E2E_SPLIT_COPIES && e->n_streams == 4) ? e->streams[3] : nullptr;
  auto issue_slice = [&](uint64_t lo, uint64_t m, cudaStream_t s, int c, bool plain_boards) -> int {
    const bool nibble = nibble_format && !plain_boards;        // this slice's boards go out 4 bits per cell
    const bool packed_wire = packed_format && !plain_boards;   // ... into the staging area, for the host threads
    if (plain_boards && nibble_format) e->h_overflow[c] = 0;   // (nothing can overflow in a plain slice)
    if (nibble) G2048_CUDA(cudaMemsetAsync(e->d_overflow + c, 0, 4, s));
    G2048_CUDA(cudaMemcpyAsync(e->d_actions + lo, actions_host + lo, m, cudaMemcpyHostToDevice, s));
    G2048StepArgs a;
    std::memset(&a, 0, sizeof a);
    a.boards = e->d_boards + 16 * lo;
    a.actions = e->d_actions + lo;
    a.rewards = e->d_rewards + lo;
    a.dones = e->d_dones + lo;
    a.illegal = o->illegal ? e->d_illegal + lo : nullptr;
    a.highest_exp = o->highest_exp ? e->d_highest + lo : nullptr;
    a.legal_mask = o->legal_mask ? e->d_mask + lo : nullptr;
    a.ep_score = e->d_ep_score + lo;
    a.ep_len = e->d_ep_len + lo;
    if (nibble) { a.boards_nibble = e->d_nibble + 8 * lo; a.nibble_overflow = e->d_overflow + c; }
    a.n = m;
    a.env_id_base = e->cfg.env_id_base + lo;
    a.seed = e->cfg.seed;
    a.step_index = e->step_index;
    a.illegal_move_reward = e->cfg.illegal_move_reward;
    a.max_tile_exp = e->cfg.max_tile_exp;
    a.flags = e->cfg.flags;
    const int rc = g2048_step(&a, s);
    if (rc) return rc;
    cudaStream_t s2 = s;
    if (aux && c < e->n_events) {
      G2048_CUDA(cudaEventRecord(e->slice_done[c], s));
      G2048_CUDA(cudaStreamWaitEvent(aux, e->slice_done[c], 0));
      s2 = aux;
    }
    if (packed_wire) {
      // packed boards into the pinned staging area; the event tells the host threads that slice c may be expanded
      G2048_CUDA(cudaMemcpyAsync(e->h_nibble + 8 * lo, e->d_nibble + 8 * lo, 8 * m, cudaMemcpyDeviceToHost, s));
      G2048_CUDA(cudaEventRecord(e->slice_done[c], s));
      G2048_CUDA(cudaMemcpyAsync(e->h_overflow + c, e->d_overflow + c, 4, cudaMemcpyDeviceToHost, s));
    } else if (nibble) {
      G2048_CUDA(cudaMemcpyAsync(o->boards + 8 * lo, e->d_nibble + 8 * lo, 8 * m, cudaMemcpyDeviceToHost, s));
      G2048_CUDA(cudaMemcpyAsync(e->h_overflow + c, e->d_overflow + c, 4, cudaMemcpyDeviceToHost, s2));
    } else {
      G2048_CUDA(cudaMemcpyAsync(o->boards + 16 * lo, e->d_boards + 16 * lo, 16 * m, cudaMemcpyDeviceToHost, s));
    }
    G2048_CUDA(cudaMemcpyAsync(o->rewards + lo, e->d_rewards + lo, 4 * m, cudaMemcpyDeviceToHost, s2));
    G2048_CUDA(cudaMemcpyAsync(o->dones + lo, e->d_dones + lo, m, cudaMemcpyDeviceToHost, s2));
    if (o->illegal) G2048_CUDA(cudaMemcpyAsync(o->illegal + lo, e->d_illegal + lo, m, cudaMemcpyDeviceToHost, s2));
    if (o->highest_exp) G2048_CUDA(cudaMemcpyAsync(o->highest_exp + lo, e->d_highest + lo, m, cudaMemcpyDeviceToHost, s2));
    if (o->legal_mask) G2048_CUDA(cudaMemcpyAsync(o->legal_mask + lo, e->d_mask + lo, m, cudaMemcpyDeviceToHost, s2));
    return G2048_OK;
  };
  if (packed_wire) {
    // the slice plan for the host threads, before anything is in flight
    int k = 0;
    for (uint64_t lo = 0; lo < n_body; ++k) {
      const uint64_t want = (k == 0 && lead) ? lead : per;
      const uint64_t m = (n_body - lo < want) ? n_body - lo : want;
      e->pool->lo[k] = lo;
      e->pool->cnt[k] = m;
      lo += m;
    }
    if (k > e->n_events) return fail(G2048_ERR_INVALID, "g2048_env_step_host: more slices than events");
    e->pool->start(e->h_nibble, o->boards, k);
  }
  int c = 0, n_packed = 0;
  for (uint64_t lo = 0; lo < n; ++c) {
    const bool in_tail = lo >= n_body;
    const uint64_t want = in_tail ? tail : ((c == 0 && lead) ? lead : per);
    const uint64_t end = in_tail ? n : n_body;
    const uint64_t m = (end - lo < want) ? end - lo : want;
    if (!in_tail) n_packed = c + 1;
    const int rc = issue_slice(lo, m, e->streams[c % (aux ? 3 : e->n_streams)], c, in_tail);
    if (rc) {
      if (packed_wire) { e->pool->abort_from(0); e->pool->finish(); }
      // Some slices of this step may already have run.  Drain the streams (no work of the failed call is left
      // in flight over the caller's host buffers) and say that the env is no longer at a step boundary: the
      // step index is NOT advanced and the boards are a mix of pre- and post-step slices — reset or
      // g2048_env_set_boards_host before stepping again.
      char first[sizeof g_err];
      std::snprintf(first, sizeof first, "%s", g_err);
      for (int i = 0; i < e->n_streams; ++i) cudaStreamSynchronize(e->streams[i]);
      return fail(rc, "g2048_env_step_host: slice %d failed (%s); env state is indeterminate (boards [0,%llu) were "
                      "stepped), step index not advanced", c, first, (unsigned long long)lo);
    }
    lo += m;
  }
  e->step_index += 1;
  if (packed_wire) {
    // hand every slice to the host threads as it lands (they expand slice k while slice k+1 is on the wire)
    for (int k = 0; k < n_packed; ++k) {
      const cudaError_t ev = cudaEventSynchronize(e->slice_done[k]);
      if (ev != cudaSuccess) {
        e->pool->abort_from(k);
        e->pool->finish();
        return cuda_fail(ev, "g2048_env_step_host: cudaEventSynchronize(slice)");
      }
      e->pool->ready[k].store(1, std::memory_order_release);
    }
  }
  const int rc = env_sync(e);
  if (packed_wire) e->pool->finish();
  if (rc == G2048_OK && nibble) {
    uint32_t total = 0;
    for (int k = 0; k < c; ++k) total += e->h_overflow[k];
    if (o->nibble_overflow) *o->nibble_overflow = total;
    if (packed_wire && total != 0) {
      // some board holds a tile >= 65536 and does not fit 4 bits per cell: this step's boards again, in full
      G2048_CUDA(cudaMemcpyAsync(o->boards, e->d_boards, n * 16, cudaMemcpyDeviceToHost, e->streams[0]));
      G2048_CUDA(cudaStreamSynchronize(e->streams[0]));
    }
  }
  return rc;
}

int g2048_env_get_boards_host(G2048Env* e, uint8_t* boards_host) {
  if (!e || !boards_host) return fail(G2048_ERR_INVALID, "g2048_env_get_boards_host: NULL argument");
  G2048_CUDA(cudaSetDevice(e->cfg.device));
  G2048_CUDA(cudaMemcpyAsync(boards_host, e->d_boards, e->cfg.n * 16, cudaMemcpyDeviceToHost, e->streams[0]));
  G2048_CUDA(cudaStreamSynchronize(e->streams[0]));
  return G2048_OK;
}

int g2048_unpack_boards_host(const uint8_t* packed, uint8_t* boards, uint64_t n) {
  if (n == 0) return G2048_OK;
  if (!packed || !boards) return fail(G2048_ERR_INVALID, "g2048_unpack_boards_host: NULL argument");
  unpack_nibble_boards(packed, boards, (size_t)n);
  return G2048_OK;
}

int g2048_env_device_ptrs(G2048Env* e, uint8_t** boards, float** rewards, uint8_t** dones) {
  if (!e) return fail(G2048_ERR_INVALID, "g2048_env_device_ptrs: env is NULL");
  if (boards) *boards = e->d_boards;
  if (rewards) *rewards = e->d_rewards;
  if (dones) *dones = e->d_dones;
  return G2048_OK;
}

uint64_t g2048_env_step_index(const G2048Env* e) { return e ? e->step_index : 0; }

}  // extern "C"
