// Ablation microbenchmark of the lean step kernel: which part of the step costs what, and how the
// launch time scales with the batch.  Includes the product's device header; not part of the product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ablate ablate.cu ; run under gpurun.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../gym-2048_b200/csrc/g2048_device.cuh"

using namespace g2048;

enum Mode { FULL = 0, MEM_ONLY = 1, NO_PHILOX = 2, NO_SPAWN = 3, NO_MOVE = 4, NO_SCORE = 5, PHILOX_ONLY = 6, NO_LOADS = 7 };

struct P {
  uint4* boards; const uint8_t* actions; float* rewards; uint8_t* dones;
  uint32_t n; uint64_t step_index; StreamKeys keys;
};

template <int MODE, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) k(const P p) {
  __shared__ Board4 s_lut[32];
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x < 32) s_lut[threadIdx.x] = one_tile_board(threadIdx.x);
  __syncthreads();
  const uint32_t n = p.n, stride = gridDim.x * THREADS;
  uint32_t i = blockIdx.x * THREADS + threadIdx.x;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (i >= n) return;
  uint4 bd; uint32_t action;
  if (MODE == NO_LOADS) { bd = make_uint4(i * 0x01010101u & 0x03030303u, 0x01020102u, i & 0x07070707u, 0x00010203u); action = i; }
  else { bd = p.boards[i]; action = p.actions[i]; }
  while (true) {
    const uint32_t i_next = i + stride;
    const bool more = i_next < n;
    uint4 bd_next = make_uint4(0, 0, 0, 0); uint32_t action_next = 0;
    if (more) {
      if (MODE == NO_LOADS) { bd_next = make_uint4(i_next * 0x01010101u & 0x03030303u, 0x01020102u, i_next & 0x07070707u, 0x00010203u); action_next = i_next; }
      else { bd_next = p.boards[i_next]; action_next = p.actions[i_next]; }
    }
    Words w;
    if (MODE == NO_PHILOX || MODE == MEM_ONLY) { w = Words{i * 0x9E3779B9u ^ (uint32_t)p.step_index, i * 0x85EBCA6Bu, i * 0xC2B2AE35u, 0}; }
    else w = words_from_pair(philox2x32_10_keys(i, p.keys));
    float reward; bool done;
    if (MODE == MEM_ONLY) {
      bd.x ^= action; reward = (float)(bd.y & 0xFF); done = (bd.z & 1u) != 0u;
    } else if (MODE == PHILOX_ONLY) {
      bd.x ^= w.w0 & 1u; bd.y ^= w.w1 & 1u; reward = (float)(w.w2 & 0xFF); done = (w.w3 & 1u) != 0u;
    } else {
      uint32_t a, b, c, d;
      float score = 0.f; bool legal = true;
      if (MODE != NO_MOVE) {
        orient(kOrientIn[action & 3u], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
        const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
        score = slide_merge(a, b, c, d);
        if (MODE == NO_SCORE) score = 0.f;
        legal = (((a ^ a0) | (b ^ b0)) | ((c ^ c0) | (d ^ d0))) != 0u;
        orient(kOrientOut[action & 3u], a, b, c, d, bd.x, bd.y, bd.z, bd.w);
      }
      uint32_t n_empty = 2;
      if (MODE != NO_SPAWN) n_empty = spawn(bd.x, bd.y, bd.z, bd.w, w.w0, legal ? 0xFFFFFFFFu : 0u);
      else bd.x ^= w.w0 & 1u;
      bool end = (n_empty == 1u) && full_board_is_dead(bd.x, bd.y, bd.z, bd.w);
      done = end || !legal;
      if (done) fresh_board(s_lut, w.w1, w.w2, bd.x, bd.y, bd.z, bd.w);
      reward = legal ? score : -1.f;
    }
    if (MODE != NO_LOADS || (bd.x == 0xdeadbeefu)) {
      p.boards[i] = bd; p.rewards[i] = reward; p.dones[i] = done ? 1 : 0;
    }
    if (!more) break;
    i = i_next; bd = bd_next; action = action_next;
  }
}

template <int MODE, int THREADS, int CTAS>
double time_mode(const char* name, uint32_t n, int sets, int steps, std::vector<uint4*>& boards, uint8_t* actions,
                 float* rewards, uint8_t* dones, bool pdl) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  P p; p.actions = actions; p.rewards = rewards; p.dones = dones; p.n = n;

  cudaLaunchConfig_t cfg = {};
  unsigned need = (n + THREADS - 1) / THREADS, cap = sms * CTAS;
  cfg.gridDim = dim3(need < cap ? need : cap); cfg.blockDim = dim3(THREADS);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    for (int t = 0; t < 20; ++t) { p.boards = boards[t % sets]; p.step_index = t; make_stream_keys(stream_key(42, t, 0, 0), t, p.keys); p.actions = actions + (n <= (1u << 20) ? (size_t)(t % 8) << 20 : 0); cudaLaunchKernelEx(&cfg, k<MODE, THREADS, CTAS>, p); }
    cudaEventRecord(e0);
    for (int t = 0; t < steps; ++t) { p.boards = boards[t % sets]; p.step_index = 100 + t; make_stream_keys(stream_key(42, 100 + t, 0, 0), 100 + t, p.keys); p.actions = actions + (n <= (1u << 20) ? (size_t)(t % 8) << 20 : 0); cudaLaunchKernelEx(&cfg, k<MODE, THREADS, CTAS>, p); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / steps;
    if (us < best) best = us;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-14s n %8u sets %2d pdl %d  %7.2f us/launch  %6.1f ps/board  %s\n", name, n, sets, (int)pdl, best, best * 1e6 / n,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  return best;
}

__global__ void init_boards(uint4* b, uint8_t* a, uint32_t n, uint32_t salt) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Words w = philox4x32_10(i, salt, 7, 9, 1, 2);
  // early/mid-game looking boards: exponents 0..3 with many zeros
  b[i] = make_uint4(w.w0 & 0x03010200u, w.w1 & 0x01030001u, w.w2 & 0x02000301u, w.w3 & 0x00020103u);
  if (a) a[i] = (uint8_t)(w.w0 >> 13);
}

int main(int argc, char** argv) {
  const uint32_t nmax = 8u << 20;
  const int max_sets = 8;
  std::vector<uint4*> boards(max_sets);
  uint8_t *actions, *dones; float* rewards;
  for (auto& b : boards) cudaMalloc(&b, (size_t)(1u << 20) * 16 * (size_t)1);
  // big sets for the scaling sweep share one allocation
  uint4* big; cudaMalloc(&big, (size_t)nmax * 16 * 2);
  cudaMalloc(&actions, nmax); cudaMalloc(&dones, nmax); cudaMalloc(&rewards, (size_t)nmax * 4);
  for (int s = 0; s < max_sets; ++s) init_boards<<<(1u << 20) / 256, 256>>>(boards[s], actions, 1u << 20, s);
  init_boards<<<2 * nmax / 256, 256>>>(big, nullptr, 2 * nmax, 99);
  init_boards<<<nmax / 256, 256>>>(big, actions, nmax, 98);
  cudaDeviceSynchronize();
  const uint32_t n = 1u << 20;
  const int steps = 2000;
#define RUN(M, T, C, sets, pdl) time_mode<M, T, C>(#M, n, sets, steps, boards, actions, rewards, dones, pdl)
  printf("== ablations at 1 Mi boards, 8 sets (HBM), 512x2, PDL\n");
  RUN(FULL, 512, 2, 8, true); RUN(MEM_ONLY, 512, 2, 8, true); RUN(NO_PHILOX, 512, 2, 8, true); RUN(NO_SPAWN, 512, 2, 8, true);
  RUN(NO_MOVE, 512, 2, 8, true); RUN(NO_SCORE, 512, 2, 8, true); RUN(PHILOX_ONLY, 512, 2, 8, true); RUN(NO_LOADS, 512, 2, 8, true);
  RUN(FULL, 512, 2, 8, true);
  if (argc > 1) return 0;
  printf("== 1 set (L2 resident)\n");
  RUN(FULL, 512, 2, 1, true); RUN(MEM_ONLY, 512, 2, 1, true);
  printf("== no PDL\n");
  RUN(FULL, 512, 2, 8, false); RUN(MEM_ONLY, 512, 2, 8, false); RUN(NO_LOADS, 512, 2, 8, false);
  printf("== scaling with n (2 big sets alternate; > L2 from 2 Mi)\n");
  for (uint32_t nn = 1u << 16; nn <= nmax; nn <<= 1) {
    std::vector<uint4*> two = {big, big + nmax};
    time_mode<FULL, 512, 2>("FULL", nn, 2, nn >= (4u << 20) ? 300 : 2000, two, actions, rewards, dones, true);
    time_mode<MEM_ONLY, 512, 2>("MEM_ONLY", nn, 2, nn >= (4u << 20) ? 300 : 2000, two, actions, rewards, dones, true);
  }
  return 0;
}
