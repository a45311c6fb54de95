// Microbenchmark (sm_100a): does the FP64 pipe of a B200 run beside the integer-multiply pipe, and how fast?
// The MSM and sumcheck kernels are bound by IMAD.WIDE (its carry-in form issues at half rate) while the FP64 pipe idles; a
// DFMA-based limb product (two fused multiply-adds give the exact high and low halves of a 51 x 51-bit product) would move
// the multiplications there. This measures the three rates that decide whether that pays:
//   mode 0: independent DFMA          mode 1: independent IMAD.WIDE.U32          mode 2: both interleaved 1:1
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rates dfma_rates.cu && ./dfma_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, double seed) {
  double d[8];
  uint64_t v[8];
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  uint32_t a = threadIdx.x * 2654435761u + 12345u, b = a ^ 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < 8; i++) { d[i] = x + i; v[i] = a + i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || MODE == 2) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(y), "d"(x));
      if (MODE == 1 || MODE == 2) {  // a multiplicand that changes every iteration (a constant product would be hoisted)
        uint32_t lo = (uint32_t)v[i];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(v[i]) : "r"(lo), "r"(b));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i] + (double)v[i];
  if (s == 0.12345) out[0] = s;
}
template <int MODE>
double run(const char *name, double ops_per_iter) {
  double *out;
  cudaMalloc(&out, 8);
  int blocks = 148 * 8;
  k<MODE><<<blocks, 256>>>(out, 1.0);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double ops = (double)blocks * 256 * ITERS * ops_per_iter;
  double rate = ops / (best * 1e-3);
  printf("%-44s %8.3f ms  %7.2f T instr/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", name, best, rate / 1e12, rate / 148 / 1.965e9);
  cudaFree(out);
  return rate;
}
int main() {
  double f = run<0>("DFMA (fma.rz.f64), 8 independent chains", 8);
  double i = run<1>("IMAD.WIDE.U32 (mad.wide.u32), 8 chains", 8);
  double b = run<2>("DFMA + IMAD.WIDE interleaved 1:1 (each)", 8);
  printf("interleaved: %.2f T DFMA/s + %.2f T IMAD.WIDE/s at once = %.2f x the DFMA-only rate, %.2f x the IMAD-only rate\n", b / 1e12, b / 1e12,
         b / f, b / i);
  return 0;
}
