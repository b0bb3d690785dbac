// g2048_internal.h — helpers shared by the translation units of libg2048.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#ifndef G2048_THREADS        // threads per CTA of every kernel in the library
#define G2048_THREADS 512
#endif

namespace g2048 {

constexpr int kThreads = G2048_THREADS;

// error reporting: set the thread-local message behind g2048_last_error() and return `code`
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int launch_check(const char* name);
// grid size for a grid-stride kernel over n items: one persistent wave over the SMs at most
unsigned grid_for(uint64_t n);
// the same with four CTAs per SM, for the low-register streaming kernels (see g2048.cu)
unsigned grid_for_streaming(uint64_t n);
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define G2048_CUDA(call)                                          \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) return ::g2048::cuda_fail(e_, #call);  \
  } while (0)

}  // namespace g2048
