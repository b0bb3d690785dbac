// Host cost of one kernel launch on this box, per API and parameter size (round 2, session 2): what bounds a shard that
// the GPU steps in ~1.3 us.  Empty kernels, 148 CTAs x 256 threads, launched back to back from one host thread.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o launch_cost launch_cost.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

template <int BYTES> struct Params { char b[BYTES]; };
template <int BYTES> __global__ void k(const Params<BYTES> p) {
  asm volatile("griddepcontrol.launch_dependents;");
  if (p.b[0] == 77 && threadIdx.x == 9999) printf("x");
}

template <int BYTES> static void run(const char* name, bool pdl, bool driver, int iters) {
  Params<BYTES> p;
  std::memset(&p, 0, sizeof p);
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (pdl) { cfg.attrs = attr; cfg.numAttrs = 1; }
  void* kargs[] = {&p};
  CUfunction fn = nullptr;
  cudaGetFuncBySymbol(&fn, (const void*)k<BYTES>);
  CUlaunchConfig dc;
  std::memset(&dc, 0, sizeof dc);
  dc.gridDimX = 148; dc.gridDimY = dc.gridDimZ = 1; dc.blockDimX = 256; dc.blockDimY = dc.blockDimZ = 1; dc.hStream = s;
  CUlaunchAttribute da[1];
  da[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
  da[0].value.programmaticStreamSerializationAllowed = 1;
  if (pdl) { dc.attrs = da; dc.numAttrs = 1; }
  for (int rep = 0; rep < 2; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) {
      if (driver) cuLaunchKernelEx(&dc, fn, kargs, nullptr);
      else cudaLaunchKernelExC(&cfg, (const void*)k<BYTES>, kargs);
    }
    auto t1 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    auto t2 = std::chrono::steady_clock::now();
    if (rep == 1)
      printf("%-28s params %4d B  %s  host %.2f us/launch  until idle %.2f us/launch\n", name, BYTES, pdl ? "PDL" : "   ",
             std::chrono::duration<double, std::micro>(t1 - t0).count() / iters,
             std::chrono::duration<double, std::micro>(t2 - t0).count() / iters);
  }
  cudaStreamDestroy(s);
}

// T host threads, one stream each, empty PDL kernels: is the ~1.9 us per launch a per-thread / per-stream limit?
static void run_threads(int T, int iters) {
  std::vector<cudaStream_t> st(T);
  for (auto& s : st) cudaStreamCreate(&s);
  auto body = [&](int t) {
    Params<320> p;
    std::memset(&p, 0, sizeof p);
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.stream = st[t];
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    void* kargs[] = {&p};
    for (int i = 0; i < iters; ++i) cudaLaunchKernelExC(&cfg, (const void*)k<320>, kargs);
  };
  for (int rep = 0; rep < 2; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(body, t);
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    cudaDeviceSynchronize();
    auto t2 = std::chrono::steady_clock::now();
    if (rep == 1)
      printf("%d threads x %d streams: host %.2f us per launch (aggregate)  until idle %.2f us per launch\n", T, T,
             std::chrono::duration<double, std::micro>(t1 - t0).count() / (iters * T),
             std::chrono::duration<double, std::micro>(t2 - t0).count() / (iters * T));
  }
  for (auto& s : st) cudaStreamDestroy(s);
}

int main() {
  cudaFree(0);
  const int N = 20000;
  run<16>("runtime cudaLaunchKernelExC", false, false, N);
  run<16>("runtime cudaLaunchKernelExC", true, false, N);
  run<320>("runtime cudaLaunchKernelExC", true, false, N);
  run<512>("runtime cudaLaunchKernelExC", true, false, N);
  run<16>("driver cuLaunchKernelEx", true, true, N);
  run<320>("driver cuLaunchKernelEx", true, true, N);
  run<512>("driver cuLaunchKernelEx", true, true, N);
  for (int T : {1, 2, 3, 4}) run_threads(T, N);
  return 0;
}
