// How fast can a kernel write its results straight into pinned host memory (no copy engine)?
// Same store pattern as the step kernel: 16 B + 4 B + 1 B per thread.  Not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint4* boards, float* rewards, uint8_t* dones, uint32_t n, uint32_t salt) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    boards[i] = make_uint4(i, salt, i ^ salt, 7u);
    rewards[i] = (float)(i & 255u);
    dones[i] = (uint8_t)(i & 1u);
  }
}
int main() {
  const uint32_t n = 1u << 20;
  uint4* hb; float* hr; uint8_t* hd;
  cudaHostAlloc(&hb, (size_t)n * 16, cudaHostAllocDefault); cudaHostAlloc(&hr, (size_t)n * 4, cudaHostAllocDefault);
  cudaHostAlloc(&hd, n, cudaHostAllocDefault);
  uint4* db; float* dr; uint8_t* dd;
  cudaMalloc(&db, (size_t)n * 16); cudaMalloc(&dr, (size_t)n * 4); cudaMalloc(&dd, n);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {148, 296, 592, 2048}) for (int threads : {256, 512}) {
    for (int w = 0; w < 3; ++w) k<<<grid, threads>>>(hb, hr, hd, n, w);
    cudaEventRecord(e0);
    for (int t = 0; t < 20; ++t) k<<<grid, threads>>>(hb, hr, hd, n, t);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("zero-copy stores grid %4d x %3d: %.3f ms per 1 Mi boards  %.1f GB/s  (%s)\n", grid, threads, ms / 20, 21.0 * n / (ms / 20 * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  }
  // reference: device kernel + 3 cudaMemcpyAsync
  cudaStream_t s; cudaStreamCreate(&s);
  for (int w = 0; w < 3; ++w) { k<<<296, 512, 0, s>>>(db, dr, dd, n, w); cudaMemcpyAsync(hb, db, (size_t)n * 16, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(hr, dr, (size_t)n * 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(hd, dd, n, cudaMemcpyDeviceToHost, s); }
  cudaEventRecord(e0, s);
  for (int t = 0; t < 20; ++t) { k<<<296, 512, 0, s>>>(db, dr, dd, n, t); cudaMemcpyAsync(hb, db, (size_t)n * 16, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(hr, dr, (size_t)n * 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(hd, dd, n, cudaMemcpyDeviceToHost, s); }
  cudaEventRecord(e1, s); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("kernel + 3 memcpy: %.3f ms per 1 Mi boards  %.1f GB/s\n", ms / 20, 21.0 * n / (ms / 20 * 1e-3) / 1e9);
  return 0;
}
