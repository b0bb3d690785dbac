// g2048_data.cu — the kernels either side of the step (sm_100a): the random policies that feed
// it actions and the transition-data transforms of the reference's training_data.py (board
// symmetries, 8x augmentation, discounted return).  Byte permutations run as PRMT networks on
// boards held in four registers, like the step kernel; everything here is HBM-bound.
#include <cuda_runtime.h>

#include "../../include/g2048.h"
#include "g2048_device.cuh"
#include "g2048_internal.h"

namespace g2048 {

// ---- random policies ------------------------------------------------------------------
// train.py:119 `random.randint(0, 3)` / the random-legal policy of BASELINE config 4, drawn
// from word 0 of the policy-tag draw stream at the step's index (its own stream: independent of
// the spawn the step makes with the step-tag words).
__global__ void __launch_bounds__(kThreads)
g2048_sample_actions_kernel(const uint8_t* legal_mask, uint8_t* actions, uint64_t n, uint64_t env_id_base,
                            uint64_t seed, uint64_t step_index) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  DrawStream draws(seed, step_index, TAG_POLICY);
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const Words w = draws.words(env_id_base + i);
    actions[i] = (uint8_t)pick_action(legal_mask ? legal_mask[i] : 15u, w.w0);
  }
}

// ---- board symmetries (training_data.py:257-279); the PRMT networks are in g2048_device.cuh ----
__device__ __forceinline__ uint4 board_hflip(uint4 b) { board_hflip(b.x, b.y, b.z, b.w); return b; }
__device__ __forceinline__ uint4 board_rot1(uint4 b) { board_rot1(b.x, b.y, b.z, b.w); return b; }

__global__ void __launch_bounds__(kThreads)
g2048_symmetry_kernel(const uint4* in, uint4* out, const uint4* next_in, uint4* next_out, const uint8_t* act_in,
                      uint8_t* act_out, uint64_t n, int hflip, int k) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    uint4 b = in[i];
    uint4 nb = next_in ? next_in[i] : make_uint4(0, 0, 0, 0);
    if (hflip) { b = board_hflip(b); nb = board_hflip(nb); }
    for (int r = 0; r < k; ++r) { b = board_rot1(b); nb = board_rot1(nb); }
    out[i] = b;
    if (next_out) next_out[i] = nb;
    if (act_in && act_out) {
      uint32_t a = act_in[i] & 3u;
      if (hflip) a = action_hflip(a);
      act_out[i] = (uint8_t)((a + (uint32_t)k) & 3u);
    }
  }
}

// augment (:281-299): one thread per source transition writes its 8 copies; for a fixed copy
// consecutive threads write consecutive rows, so every store instruction is coalesced.
__global__ void __launch_bounds__(kThreads)
g2048_augment_kernel(const uint4* boards, const uint4* next_boards, const uint8_t* actions, const float* rewards,
                     const uint8_t* dones, uint64_t n, uint4* boards_out, uint4* next_out, uint8_t* actions_out,
                     float* rewards_out, uint8_t* dones_out) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const uint4 b0 = boards[i], n0 = next_boards[i];
    const uint32_t a0 = actions[i] & 3u;
    const float r = rewards[i];
    const uint8_t d = dones[i];
    uint4 b[2] = {b0, board_hflip(b0)};
    uint4 nb[2] = {n0, board_hflip(n0)};
    const uint32_t a[2] = {a0, action_hflip(a0)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t o = (uint64_t)(2 * k + h) * n + i;
        boards_out[o] = b[h];
        next_out[o] = nb[h];
        actions_out[o] = (uint8_t)((a[h] + (uint32_t)k) & 3u);
        rewards_out[o] = r;
        dones_out[o] = d;
        b[h] = board_rot1(b[h]);
        nb[h] = board_rot1(nb[h]);
      }
    }
  }
}

// get_discounted_return (:104-124).  Rows are one chain in game order; `done` cuts it.  A thread
// that sits on a segment end (done row, or the last row) walks its segment backwards in the
// reference's evaluation order; __dmul_rn/__dadd_rn keep the two roundings of `r + gamma * prev`.
__global__ void __launch_bounds__(kThreads)
g2048_discounted_return_kernel(const float* rewards, const uint8_t* dones, double* returns, uint64_t n,
                               double gamma) {
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    if (!(dones[i] || i == n - 1)) continue;
    double g = (double)rewards[i];
    returns[i] = g;
    for (uint64_t j = i; j-- > 0 && !dones[j];) {
      // the reference skips the add when the running value is falsy (0.0); r + gamma*0 == r
      g = __dadd_rn((double)rewards[j], __dmul_rn(gamma, g));
      returns[j] = g;
    }
  }
}

// GAE(lambda) over a time-major [T,n] rollout (SB3 RolloutBuffer.compute_returns_and_advantage).
// Thread i owns env i and walks t = T-1..0; at every t the warp reads/writes consecutive floats.
__global__ void __launch_bounds__(kThreads)
g2048_gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                 const uint8_t* __restrict__ episode_starts, const float* __restrict__ last_values,
                 const uint8_t* __restrict__ last_dones, float* __restrict__ advantages, float* __restrict__ returns,
                 uint64_t T, uint64_t n, float gamma,
                 float gl) {                       // gl = float(gamma * gae_lambda), product taken in double
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    float next_value = last_values[i];
    float next_non_terminal = last_dones[i] ? 0.f : 1.f;
    float gae = 0.f;
    // The recurrence is serial in t but its loads are not: unrolled (and the arrays declared non-aliasing), the
    // loads of eight time steps are in flight while one step's arithmetic runs (180 -> 127 us for 256 x 65,536;
    // 128-thread CTAs with a 16-deep unroll were no faster).
#pragma unroll 8
    for (uint64_t t = T; t-- > 0;) {
      const uint64_t o = t * n + i;
      const float v = values[o];
      const float delta = __fsub_rn(__fadd_rn(rewards[o], __fmul_rn(__fmul_rn(gamma, next_value), next_non_terminal)), v);
      gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, next_non_terminal), gae));
      advantages[o] = gae;
      returns[o] = __fadd_rn(gae, v);
      next_value = v;
      next_non_terminal = episode_starts[o] ? 0.f : 1.f;
    }
  }
}

}  // namespace g2048

using namespace g2048;

extern "C" {

int g2048_sample_actions(const uint8_t* legal_mask, uint8_t* actions, uint64_t n, uint64_t env_id_base,
                         uint64_t seed, uint64_t step_index, void* stream) {
  if (n == 0) return G2048_OK;
  if (!actions) return fail(G2048_ERR_INVALID, "g2048_sample_actions: actions is NULL");
  g2048_sample_actions_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      legal_mask, actions, n, env_id_base, seed, step_index);
  return launch_check("g2048_sample_actions_kernel");
}

int g2048_symmetry(const uint8_t* boards_in, uint8_t* boards_out, const uint8_t* next_in, uint8_t* next_out,
                   const uint8_t* actions_in, uint8_t* actions_out, uint64_t n, int hflip, int k, void* stream) {
  if (n == 0) return G2048_OK;
  if (!boards_in || !boards_out) return fail(G2048_ERR_INVALID, "g2048_symmetry: boards_in and boards_out are required");
  if ((next_in == nullptr) != (next_out == nullptr) || (actions_in == nullptr) != (actions_out == nullptr))
    return fail(G2048_ERR_INVALID, "g2048_symmetry: next_in/next_out and actions_in/actions_out come in pairs");
  if (!aligned16(boards_in) || !aligned16(boards_out) || !aligned16(next_in) || !aligned16(next_out))
    return fail(G2048_ERR_ALIGN, "g2048_symmetry: boards must be 16-byte aligned");
  g2048_symmetry_kernel<<<grid_for_streaming(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(boards_in), reinterpret_cast<uint4*>(boards_out),
      reinterpret_cast<const uint4*>(next_in), reinterpret_cast<uint4*>(next_out), actions_in, actions_out, n,
      hflip ? 1 : 0, ((k % 4) + 4) % 4);
  return launch_check("g2048_symmetry_kernel");
}

int g2048_augment(const uint8_t* boards, const uint8_t* next_boards, const uint8_t* actions, const float* rewards,
                  const uint8_t* dones, uint64_t n, uint8_t* boards_out, uint8_t* next_boards_out,
                  uint8_t* actions_out, float* rewards_out, uint8_t* dones_out, void* stream) {
  if (n == 0) return G2048_OK;
  if (!boards || !next_boards || !actions || !rewards || !dones || !boards_out || !next_boards_out || !actions_out ||
      !rewards_out || !dones_out)
    return fail(G2048_ERR_INVALID, "g2048_augment: NULL pointer");
  if (!aligned16(boards) || !aligned16(next_boards) || !aligned16(boards_out) || !aligned16(next_boards_out))
    return fail(G2048_ERR_ALIGN, "g2048_augment: boards must be 16-byte aligned");
  g2048_augment_kernel<<<grid_for_streaming(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(boards), reinterpret_cast<const uint4*>(next_boards), actions, rewards, dones, n,
      reinterpret_cast<uint4*>(boards_out), reinterpret_cast<uint4*>(next_boards_out), actions_out, rewards_out,
      dones_out);
  return launch_check("g2048_augment_kernel");
}

int g2048_discounted_return(const float* rewards, const uint8_t* dones, double* returns, uint64_t n, double gamma,
                            void* stream) {
  if (n == 0) return G2048_OK;
  if (!rewards || !dones || !returns) return fail(G2048_ERR_INVALID, "g2048_discounted_return: NULL pointer");
  g2048_discounted_return_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      rewards, dones, returns, n, gamma);
  return launch_check("g2048_discounted_return_kernel");
}

int g2048_gae(const float* rewards, const float* values, const uint8_t* episode_starts, const float* last_values,
              const uint8_t* last_dones, float* advantages, float* returns, uint64_t T, uint64_t n, double gamma,
              double gae_lambda, void* stream) {
  if (n == 0 || T == 0) return G2048_OK;
  if (!rewards || !values || !episode_starts || !last_values || !last_dones || !advantages || !returns)
    return fail(G2048_ERR_INVALID, "g2048_gae: NULL pointer");
  g2048_gae_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      rewards, values, episode_starts, last_values, last_dones, advantages, returns, T, n, (float)gamma,
      (float)(gamma * gae_lambda));
  return launch_check("g2048_gae_kernel");
}

}  // extern "C"
