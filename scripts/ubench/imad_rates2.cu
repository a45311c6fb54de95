// Microbenchmark (sm_100a): how fast can a B200 do 32 x 32 + 64 -> 64 multiply-accumulates, and in which form?
// The round-1 benchmark (imad_rates.cu modes 0 / 4, and the library's k_imad_peak) fed loop-invariant operands to
// `mad.wide.u32`: ptxas hoisted the product out of the loop and what was timed was a stream of IADD3 / IADD3.X pairs - no
// multiplies at all. Here the multiplicand changes every iteration, and the SASS of each mode is checked (cuobjdump) to be
// what its name says:
//   mode 0  IMAD.WIDE.U32 Rd, Ra, Rb, RZ          plain product, xor-ed into the accumulator (LOP3 on the ALU pipe)
//   mode 1  IMAD.WIDE.U32 Rd, Ra, Rb, Rd          single-instruction multiply-accumulate (carry-flag pair form)
//   mode 2  IMAD.WIDE.U32 Rd, Ra, Rb, RZ + IADD3 / IADD3.X   product on the multiply pipe, 64-bit add on the ALU pipe
//   mode 3  one in three accumulations in the multiplier, two on the ALU pipe
//   mode 4  IADD3 / IADD3.X pairs only (what round 1 measured)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
  uint32_t x = threadIdx.x * 2654435761u + seed, y = x ^ 0x9e3779b9u;
  uint64_t v[12];
  uint32_t av[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { v[i] = x + i; av[i] = x * (2 * i + 3); }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
      uint32_t a = av[i];  // fixed per thread; the other factor, y, changes every iteration (one add per 12 products)
      if (MODE == 0) {
        uint64_t p;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(y));
        v[i] ^= p;
      } else if (MODE == 1 || (MODE == 3 && i % 3 == 0)) {
        uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(y));
        v[i] = ((uint64_t)hi << 32) | lo;
      } else if (MODE == 2 || MODE == 3) {
        uint64_t p;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(y));
        uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"((uint32_t)p), "r"((uint32_t)(p >> 32)));
        v[i] = ((uint64_t)hi << 32) | lo;
      } else {
        uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(y));
        v[i] = ((uint64_t)hi << 32) | lo;
      }
    }
    y += 0x9e3779b1u;
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) s ^= v[i];
  if (s == 0x123456789abcdefull) out[0] = (uint32_t)s;
}
template <int MODE>
void run(const char *name) {
  uint32_t *out;
  cudaMalloc(&out, 4);
  int blocks = 148 * 8;
  k<MODE><<<blocks, 256>>>(out, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double ops = (double)blocks * 256 * ITERS * 12;
  printf("%-72s %8.3f ms  %7.2f Tops/s  (%5.2f lanes per clk per SM at 1.965 GHz)\n", name, best, ops / best / 1e9, ops / (best * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("0: IMAD.WIDE.U32 product, no addend (xor into the accumulator)");
  run<1>("1: IMAD.WIDE.U32 with 64-bit addend (single-instruction MAC)");
  run<2>("2: IMAD.WIDE.U32 product + IADD3/IADD3.X on the ALU pipe");
  run<3>("3: one in three accumulations in the multiplier, two on the ALU");
  run<4>("4: IADD3/IADD3.X pairs only (what round 1 called the IMAD peak)");
  return 0;
}
