// Microbenchmark: issue rates of the integer-multiply forms used by the field arithmetic (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
  uint32_t x = threadIdx.x * 2654435761u + seed, y = x ^ 0x9e3779b9u;
  uint32_t a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = x + i;
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {  // independent mad.wide (no carries)
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint64_t v = ((uint64_t)a[i + 1] << 32) | a[i];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(v) : "r"(x), "r"(y));
        a[i] = (uint32_t)v; a[i + 1] = (uint32_t)(v >> 32);
      }
    } else if (MODE == 1) {  // one carry chain of 8 IMAD.WIDE.X (mad.lo.cc/madc.hi.cc pairs)
      asm volatile("mad.lo.cc.u32 %0, %16, %17, %0; madc.hi.cc.u32 %1, %16, %17, %1;"
          "madc.lo.cc.u32 %2, %16, %17, %2; madc.hi.cc.u32 %3, %16, %17, %3;"
          "madc.lo.cc.u32 %4, %16, %17, %4; madc.hi.cc.u32 %5, %16, %17, %5;"
          "madc.lo.cc.u32 %6, %16, %17, %6; madc.hi.cc.u32 %7, %16, %17, %7;"
          "madc.lo.cc.u32 %8, %16, %17, %8; madc.hi.cc.u32 %9, %16, %17, %9;"
          "madc.lo.cc.u32 %10, %16, %17, %10; madc.hi.cc.u32 %11, %16, %17, %11;"
          "madc.lo.cc.u32 %12, %16, %17, %12; madc.hi.cc.u32 %13, %16, %17, %13;"
          "madc.lo.cc.u32 %14, %16, %17, %14; madc.hi.u32 %15, %16, %17, %15;"
          : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
            "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
          : "r"(x), "r"(y));
    } else if (MODE == 2) {  // two independent chains of 4
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
          "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
          "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
          "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
          : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) : "r"(x), "r"(y));
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
          "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
          "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
          "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
          : "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]) : "r"(x), "r"(y));
    } else if (MODE == 3) {  // 32-bit mad.lo (IMAD) independent
#pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
    } else if (MODE == 4) {  // mad.wide independent, 16 accumulators of 64 bit (8 per iter x2)
      uint64_t v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = ((uint64_t)a[2 * i + 1] << 32) | a[2 * i];
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(v[i]) : "r"(x + r), "r"(y));
#pragma unroll
      for (int i = 0; i < 8; i++) { a[2 * i] = (uint32_t)v[i]; a[2 * i + 1] = (uint32_t)(v[i] >> 32); }
    } else if (MODE == 5) {  // add.cc chain of 16 (IADD3.X)
      asm volatile("add.cc.u32 %0, %0, %16; addc.cc.u32 %1, %1, %16; addc.cc.u32 %2, %2, %16; addc.cc.u32 %3, %3, %16;"
          "addc.cc.u32 %4, %4, %16; addc.cc.u32 %5, %5, %16; addc.cc.u32 %6, %6, %16; addc.cc.u32 %7, %7, %16;"
          "addc.cc.u32 %8, %8, %16; addc.cc.u32 %9, %9, %16; addc.cc.u32 %10, %10, %16; addc.cc.u32 %11, %11, %16;"
          "addc.cc.u32 %12, %12, %16; addc.cc.u32 %13, %13, %16; addc.cc.u32 %14, %14, %16; addc.u32 %15, %15, %16;"
          : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
            "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
          : "r"(x));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= a[i];
  if (s == 0x12345678u) out[0] = s;
}
template <int MODE>
void run(const char *name, double ops_per_iter) {
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
  double ops = (double)blocks * 256 * ITERS * ops_per_iter;
  printf("%-44s %8.3f ms  %7.2f Tops/s  (%5.2f per clk per SM at 1.92 GHz)\n", name, best, ops / best / 1e9, ops / (best * 1e-3) / 148 / 1.92e9);
  cudaFree(out);
}
int main() {
  run<0>("mad.wide.u32 independent (8/iter)", 8);
  run<4>("mad.wide.u32 independent (16/iter)", 16);
  run<1>("IMAD.WIDE.X one carry chain of 8", 8);
  run<2>("IMAD.WIDE.X two carry chains of 4", 8);
  run<3>("mad.lo.u32 independent (16/iter)", 16);
  run<5>("add.cc chain of 16", 16);
  return 0;
}
