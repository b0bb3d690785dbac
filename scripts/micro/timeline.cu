// Where does a 1 Mi-board step launch spend its time?  The lean step loop of the product (same device header,
// same launch shape, PDL) with %globaltimer stamps taken by thread 0 of every CTA: kernel entry, release of
// griddepcontrol.wait, end of the first board, end of the loop.  Not part of the product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o timeline timeline.cu ; run under gpurun.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../gym-2048_b200/csrc/g2048_device.cuh"

using namespace g2048;

struct P {
  uint4* boards; const uint8_t* actions; float* rewards; uint8_t* dones;
  uint32_t n; StreamKeys keys; unsigned long long* trace;   // [grid][4]
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

template <int THREADS, int CTAS, bool TRACE>
__global__ void __launch_bounds__(THREADS, CTAS) k(const P p) {
  __shared__ Board4 s_lut[32];
  __shared__ Sel4 s_sel[8];
  asm volatile("griddepcontrol.launch_dependents;");
  unsigned long long t_entry = 0, t_wait = 0, t_first = 0;
  if (TRACE && (threadIdx.x & 31) == 0) t_entry = gtime();
  if (threadIdx.x >= 32 && threadIdx.x < 40) { const uint32_t q = threadIdx.x - 32; s_sel[q] = (q < 4) ? kOrientIn[q] : kOrientOut[q - 4]; }
  if (threadIdx.x < 32) s_lut[threadIdx.x] = one_tile_board(threadIdx.x);
  __syncthreads();
  const uint32_t n = p.n, stride = gridDim.x * THREADS;
  uint32_t i = blockIdx.x * THREADS + threadIdx.x;
#ifndef NO_PRE_PREFETCH
  if (i < n) {      // the product's prologue prefetch (g2048.cu): first boards into the L2 while the previous launch drains
    if ((threadIdx.x & 7u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.boards + i));
    if ((threadIdx.x & 31u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.actions + i));
  }
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (TRACE && (threadIdx.x & 31) == 0) t_wait = gtime();
  if (i >= n) return;
  uint4 bd = p.boards[i]; uint32_t action = p.actions[i];
  bool first = true;
  while (true) {
    const uint32_t i_next = i + stride;
    const bool more = i_next < n;
    const uint32_t act = action & 3u;
    uint32_t a, b, c, d;
    orient(s_sel[act], bd.x, bd.y, bd.z, bd.w, a, b, c, d);
    const Sel4 so = s_sel[4u + act];
    if (more) { bd = p.boards[i_next]; action = p.actions[i_next]; }
    const Words w = words_from_pair(philox2x32_10_keys(i, p.keys));
    uint4 o4;
    const StepOut o = step_oriented(s_lut, a, b, c, d, so, w, 0u, false, true, o4.x, o4.y, o4.z, o4.w);
    p.boards[i] = o4; p.rewards[i] = o.legal ? o.score : -1.f; p.dones[i] = o.done ? 1 : 0;
    if (TRACE && first && (threadIdx.x & 31) == 0) { t_first = gtime(); first = false; }
    if (!more) break;
    i = i_next;
  }
  if (TRACE && (threadIdx.x & 31) == 0) {
    unsigned long long* t = p.trace + 4ull * (blockIdx.x * (THREADS / 32) + threadIdx.x / 32);
    uint32_t smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    t[0] = smid; t[1] = t_wait; t[2] = t_first; t[3] = gtime();
  }
}

__global__ void init_boards(uint4* b, uint8_t* a, uint32_t n, uint32_t salt) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Words w = philox4x32_10(i, salt, 7, 9, 1, 2);
  b[i] = make_uint4(w.w0 & 0x03010200u, w.w1 & 0x01030001u, w.w2 & 0x02000301u, w.w3 & 0x00020103u);
  if (a) a[i] = (uint8_t)(w.w0 >> 13);
}

int main() {
  const uint32_t n = 1u << 20; const int sets = 8, T = 1024, C = 1;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const unsigned grid = sms * C;
  std::vector<uint4*> boards(sets);
  uint8_t *actions, *dones; float* rewards;
  for (auto& b : boards) cudaMalloc(&b, (size_t)n * 16);
  cudaMalloc(&actions, (size_t)n * 8); cudaMalloc(&dones, n); cudaMalloc(&rewards, (size_t)n * 4);
  for (int s = 0; s < sets; ++s) init_boards<<<n / 256, 256>>>(boards[s], actions + (size_t)s * n, n, s);
  const int L = 40;                                 // traced launches
  unsigned long long* trace; cudaMalloc(&trace, (size_t)L * grid * (T / 32) * 4 * 8);
  cudaMemset(trace, 0, (size_t)L * grid * (T / 32) * 4 * 8);
  cudaDeviceSynchronize();
  P p; p.actions = actions; p.rewards = rewards; p.dones = dones; p.n = n;
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T);
  cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = attr; cfg.numAttrs = 1;
  auto launch = [&](int t, bool tr, int slot) {
    p.boards = boards[t % sets]; p.actions = actions + (size_t)(t % 8) * n;
    make_stream_keys(stream_key(42, t, 0, 0), t, p.keys);
    p.trace = trace + (size_t)slot * grid * (T / 32) * 4;
    if (tr) cudaLaunchKernelEx(&cfg, k<1024, 1, true>, p); else cudaLaunchKernelEx(&cfg, k<1024, 1, false>, p);
  };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int t = 0; t < 200; ++t) launch(t, false, 0);
  cudaEventRecord(e0);
  for (int t = 0; t < 2000; ++t) launch(t, false, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("untraced: %.2f us/launch\n", ms * 1e3 / 2000);
  for (int t = 0; t < 100; ++t) launch(t, true, 0);
  cudaEventRecord(e0);
  for (int t = 0; t < L; ++t) launch(t, true, t);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("traced:   %.2f us/launch\n", ms * 1e3 / L);
  const unsigned WPC = T / 32, W = grid * WPC;
  std::vector<unsigned long long> h((size_t)L * W * 4);
  cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost);
  // distribution of warp end times, by SM, by warp index in the CTA
  for (int l = L - 3; l < L - 1; ++l) {
    const unsigned long long* t = &h[(size_t)l * W * 4];
    unsigned long long base = ~0ull;
    for (unsigned w = 0; w < W; ++w) base = std::min(base, t[4 * w + 1]);
    const unsigned long long* tn = &h[(size_t)(l + 1) * W * 4];
    unsigned long long next_rel = ~0ull;
    for (unsigned w = 0; w < W; ++w) next_rel = std::min(next_rel, tn[4 * w + 1]);
    std::vector<double> sm_end(sms, 0), sm_first(sms, 0);
    std::vector<double> ends, firsts;
    double by_warp[32] = {0}, by_warp_first[32] = {0};
    for (unsigned w = 0; w < W; ++w) {
      const double e = (double)t[4 * w + 3] - (double)base, f = (double)t[4 * w + 2] - (double)base;
      ends.push_back(e); firsts.push_back(f);
      const unsigned sm = (unsigned)t[4 * w + 0];
      sm_end[sm] = std::max(sm_end[sm], e); sm_first[sm] = std::max(sm_first[sm], f);
      by_warp[w % WPC] += e / grid; by_warp_first[w % WPC] += f / grid;
    }
    std::sort(ends.begin(), ends.end()); std::sort(firsts.begin(), firsts.end());
    printf("launch %d: next release at %.0f ns after this release; this launch's last warp ended at %.0f ns\n", l,
           (double)next_rel - (double)base, ends[W - 1]);
    printf("  warp first-board-done percentiles 0/10/50/90/100: %.0f %.0f %.0f %.0f %.0f\n", firsts[0], firsts[W / 10], firsts[W / 2], firsts[W * 9 / 10], firsts[W - 1]);
    printf("  warp loop-end percentiles        0/10/50/90/100: %.0f %.0f %.0f %.0f %.0f\n", ends[0], ends[W / 10], ends[W / 2], ends[W * 9 / 10], ends[W - 1]);
    std::vector<double> se = sm_end; std::sort(se.begin(), se.end());
    printf("  per-SM last warp end percentiles 0/10/50/90/100: %.0f %.0f %.0f %.0f %.0f\n", se[0], se[sms / 10], se[sms / 2], se[sms * 9 / 10], se[sms - 1]);
    printf("  mean end by warp index in CTA:");
    for (unsigned q = 0; q < WPC; ++q) printf(" %.0f", by_warp[q]);
    if (l == L - 2) {
      for (unsigned smq = 0; smq < 2; ++smq) {
        printf("\n  SM %u warps (cta.warp first end):", smq);
        for (unsigned w = 0; w < W; ++w) if ((unsigned)t[4 * w + 0] == smq)
          printf(" %u.%u %.1f %.1f |", w / WPC, w % WPC, ((double)t[4 * w + 2] - (double)base) / 1000., ((double)t[4 * w + 3] - (double)base) / 1000.);
      }
    }
    printf("\n  mean first-done by warp index:");
    for (unsigned q = 0; q < WPC; ++q) printf(" %.0f", by_warp_first[q]);
    printf("\n");
  }
  return 0;
}
