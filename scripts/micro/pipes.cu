// Microbenchmark: sustained issue rate of ALU (LOP3/PRMT) and FMA-heavy (IMAD) streams on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 4
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t one, uint32_t seed, int iters) {
  uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { a[c] = seed + threadIdx.x * 7 + c; b[c] = seed ^ (c * 0x9E3779B9u); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        if (MODE == 0) {          // ALU only: 2 LOP3 (3-input) per chain
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(one));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[c]) : "r"(a[c]), "r"(one));
        } else if (MODE == 1) {   // IMAD only
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(one), "r"(b[c]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(one), "r"(a[c]));
        } else if (MODE == 2) {   // 1:1 LOP3 : IMAD
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(one));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(one), "r"(a[c]));
        } else if (MODE == 3) {   // PRMT only
          asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a[c]) : "r"(b[c]));
          asm volatile("prmt.b32 %0, %0, %1, 0x7362;" : "+r"(b[c]) : "r"(a[c]));
        } else if (MODE == 4) {   // IMAD.WIDE-like: mul.hi
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(one));
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(a[c]), "r"(one));
        } else if (MODE == 5) {   // 2:1 LOP3 : IMAD
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(one));
          asm volatile("prmt.b32 %0, %0, %1, 0x7362;" : "+r"(b[c]) : "r"(a[c]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(one), "r"(b[c]));
        } else if (MODE == 6) {   // 1:1 LOP3 : FADD (fmalite-capable)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(one));
          float f = __uint_as_float(b[c]); f = f + __uint_as_float(a[c]); b[c] = __float_as_uint(f);
        } else if (MODE == 7) {   // 1:1:1 LOP3 : IMAD : FADD
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(one));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(one), "r"(a[c]));
          float f = __uint_as_float(b[c]); f = f + __uint_as_float(a[c]); b[c] = __float_as_uint(f);
        }
        else if (MODE == 8) {   // philox-like: IMAD.WIDE + LOP3
          uint32_t hi, lo;
          asm volatile("mul.hi.u32 %0, %2, %3; mul.lo.u32 %1, %2, %3;" : "=r"(hi), "=r"(lo) : "r"(a[c]), "r"(0xD2511F53u));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[c]) : "r"(hi), "r"(b[c]), "r"(one));
          b[c] = lo;
        }
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) r ^= a[c] + b[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, int per_iter_instr, int ctas_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = sms * ctas_per_sm, iters = 2000;
  uint32_t* out; cudaMalloc(&out, grid * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, 1, 3, 10);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 1, 3, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double warp_instr = (double)grid * 8 * iters * per_iter_instr;          // 8 warps per CTA
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s warps/SMSP %2d  IPC/SMSP %.3f\n", name, ctas_per_sm * 2, warp_instr / cycles / (sms * 4));
  cudaFree(out);
}

int main() {
  for (int c : {2, 4, 8}) {
    run<0>("LOP3 only", 8 * CHAINS * 2, c);
    run<3>("PRMT only", 8 * CHAINS * 2, c);
    run<1>("IMAD only", 8 * CHAINS * 2, c);
    run<4>("IMAD.HI only", 8 * CHAINS * 2, c);
    run<2>("LOP3:IMAD 1:1", 8 * CHAINS * 2, c);
    run<5>("LOP3:IMAD 2:1", 8 * CHAINS * 3, c);
    run<6>("LOP3:FADD 1:1", 8 * CHAINS * 2, c);
    run<7>("LOP3:IMAD:FADD 1:1:1", 8 * CHAINS * 3, c);
    run<8>("WIDE:LOP3 1:1 (philox)", 8 * CHAINS * 2, c);
  }
  return 0;
}
