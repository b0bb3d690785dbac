// Microbenchmark 2: which pipe do the less common integer instructions use on sm_100a?  Each mode runs one
// instruction kind alone and paired 1:1 with LOP3 (ALU pipe) and with IMAD (FMA-heavy pipe): an instruction that
// shares a pipe with its partner halves the pair's IPC, one on another pipe does not.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 4
enum Kind { LOP3, IMAD, VABS, VSAD, POPC, SHFV, IADD3, FADD, FFMA, HADD2, I2FP, DP4A, PRMTV, SELP, IMADHI, FLOP, BREV, LEAV, FMNMX, IMNMX, PRMTR, LOP3I, IMADW, SHFC, ISETPV, VIADDV, KINDS };
static const char* kNames[] = {"LOP3", "IMAD", "VABSDIFF4", "VABSDIFF4.ACC", "POPC", "SHF", "IADD3", "FADD", "FFMA", "HADD2", "I2FP", "IDP.4A", "PRMT", "ISETP+SEL", "IMAD.HI", "FLO", "BREV", "LEA", "FMNMX", "VIMNMX", "PRMT (reg selector)", "LOP3 (imm operand)", "IMAD.WIDE", "SHF (const)", "ISETP only", "VIADD (imm)"};

template <int K> __device__ __forceinline__ void op(uint32_t& a, uint32_t& b, uint32_t one) {
  if (K == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(one));
  else if (K == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(one), "r"(b));
  else if (K == VABS) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(0));
  else if (K == VSAD) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(one));
  else if (K == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(a));
  else if (K == SHFV) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(one));
  else if (K == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
  else if (K == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&a) : "f"(*(float*)&b));
  else if (K == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a) : "f"(*(float*)&b), "f"(*(float*)&one));
  else if (K == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  else if (K == I2FP) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(*(float*)&a) : "r"(a));
  else if (K == DP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(one));
  else if (K == PRMTV) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
  else if (K == SELP) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(a) : "r"(b), "r"(one));
  else if (K == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(one));
  else if (K == FLOP) asm volatile("bfind.u32 %0, %0;" : "+r"(a));
  else if (K == BREV) asm volatile("brev.b32 %0, %0;" : "+r"(a));
  else if (K == LEAV) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(a) : "r"(b));
  else if (K == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float*)&a) : "f"(*(float*)&b));
  else if (K == IMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
  else if (K == PRMTR) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(one));
  else if (K == LOP3I) asm volatile("lop3.b32 %0, %0, %1, 0x7f7f7f7f, 0x96;" : "+r"(a) : "r"(b));
  else if (K == IMADW) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a), "r"(0xD256D193u)); a = (uint32_t)(w >> 32) ^ (uint32_t)w; }
  else if (K == SHFC) asm volatile("shr.u32 %0, %0, 7;" : "+r"(a));
  else if (K == ISETPV) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, 1; }" : "+r"(a) : "r"(b));
  else if (K == VIADDV) asm volatile("add.u32 %0, %0, 0x7f7f7f7f;" : "+r"(a));
}

template <int K1, int K2>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t one, uint32_t seed, int iters) {
  uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { a[c] = seed + threadIdx.x * 7 + c; b[c] = seed ^ (c * 0x9E3779B9u); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        op<K1>(a[c], b[c], one);
        if (K2 >= 0) op<(K2 >= 0 ? K2 : 0)>(b[c], a[c], one);
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) r ^= a[c] + b[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int K1, int K2> double run(int ctas_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = sms * ctas_per_sm, iters = 1500;
  uint32_t* out; cudaMalloc(&out, grid * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<K1, K2><<<grid, 256>>>(out, 1, 3, 10);
  cudaEventRecord(e0);
  k<K1, K2><<<grid, 256>>>(out, 1, 3, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int n_ops = (K2 >= 0 ? 2 : 1) * ((K1 == SELP || K1 == LEAV) ? 1 : 1);
  double warp_instr = (double)grid * 8 * iters * 8 * CHAINS * n_ops;     // PTX-level ops (SELP/LEA may be 2 SASS)
  double cycles = ms * 1e-3 * clk * 1e3;
  cudaFree(out);
  return warp_instr / cycles / (sms * 4);
}

// dependent-issue latency: ONE warp per SM sub-partition, ONE dependent chain
template <int K>
__global__ void __launch_bounds__(128) klat(uint32_t* out, uint32_t one, uint32_t seed, int iters) {
  uint32_t a = seed + threadIdx.x * 7, b = seed ^ 0x9E3779B9u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 32; ++u) op<K>(a, b, one);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
}
template <int K> double latency() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4000;
  uint32_t* out; cudaMalloc(&out, sms * 128 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  klat<K><<<sms, 128>>>(out, 1, 3, 10);
  cudaEventRecord(e0);
  klat<K><<<sms, 128>>>(out, 1, 3, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaFree(out);
  return ms * 1e-3 * clk * 1e3 / ((double)iters * 32);
}

template <int K> void row() {
  printf("%-20s alone %.3f   with LOP3 %.3f   with IMAD %.3f   with FADD %.3f   dependent-issue latency %.1f cycles\n", kNames[K],
         run<K, -1>(4), run<K, LOP3>(4), run<K, IMAD>(4), run<K, FADD>(4), latency<K>());
}

int main() {
  printf("PTX ops per cycle per SM sub-partition, 8 warps/SMSP, 4 independent chains per thread\n");
  row<LOP3>(); row<IMAD>(); row<IMADHI>(); row<PRMTV>(); row<SHFV>(); row<IADD3>(); row<VABS>(); row<VSAD>(); row<POPC>();
  row<FLOP>(); row<BREV>(); row<SELP>(); row<LEAV>(); row<IMNMX>(); row<FADD>(); row<FFMA>(); row<FMNMX>(); row<HADD2>();
  row<I2FP>(); row<DP4A>(); row<PRMTR>(); row<LOP3I>(); row<IMADW>(); row<SHFC>(); row<ISETPV>(); row<VIADDV>();
  return 0;
}
