// Microbenchmark (sm_100a): LATENCY of one F_p multiplication for a lone warp — what bounds the Horner / encoding kernel, the
// bullet-round MSMs and every other kernel that is a chain of dependent multiplications with nothing else to run.
//   variant 0: limb::mul_8x8 as shipped (every row accumulates into the same even/odd accumulator pair: 32 dependent IMAD.WIDE.X)
//   variant 1: rows split over 2 independent accumulator pairs, merged at the end
//   variant 2: rows split over 4 independent accumulator pairs
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../vpin_b200/csrc -o mul_latency mul_latency.cu && ./mul_latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ed.cuh"
using namespace vpin;

// t[0..16) += s[0..n) << (32 * off), carry propagated to the top
__device__ __forceinline__ void add_shifted(uint32_t *t, const uint32_t *s, int n, int off) {
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    if (k >= off) {
      c += (uint64_t)t[k] + (k - off < n ? s[k - off] : 0u);
      t[k] = (uint32_t)c;
      c >>= 32;
    }
  }
}
// product of a (8 limbs) with b[lo .. lo + cnt) as an (8 + cnt)-limb number, own even/odd accumulators
template <int CNT>
__device__ __forceinline__ void mul_8xN(uint32_t *out, const uint32_t *a, const uint32_t *b) {
  uint32_t ev[8 + CNT + 2], od[8 + CNT + 2];
#pragma unroll
  for (int k = 0; k < 8 + CNT + 2; k++) ev[k] = od[k] = 0;
  limb::mul_row(ev, a, b[0]);
  limb::mul_row(od, a + 1, b[0]);
#pragma unroll
  for (int i = 1; i < CNT; i++) {
    if (i & 1) {
      limb::mad_row(od + i - 1, a, b[i], od[i + 7]);
      limb::mad_row(ev + i + 1, a + 1, b[i], ev[i + 9]);
    } else {
      limb::mad_row(ev + i, a, b[i], ev[i + 8]);
      limb::mad_row(od + i, a + 1, b[i], od[i + 8]);
    }
  }
  uint64_t c = 0;
  out[0] = ev[0];
#pragma unroll
  for (int k = 1; k < 8 + CNT; k++) {
    c += (uint64_t)ev[k] + od[k - 1];
    out[k] = (uint32_t)c;
    c >>= 32;
  }
}
template <int GROUPS>
__device__ __forceinline__ fp_t fp_mul_ilp(const fp_t &a, const fp_t &b) {
  constexpr int CNT = 8 / GROUPS;
  uint32_t t[16];
  uint32_t part[GROUPS][8 + CNT];
#pragma unroll
  for (int g = 0; g < GROUPS; g++) mul_8xN<CNT>(part[g], a.v, b.v + g * CNT);
#pragma unroll
  for (int k = 0; k < 16; k++) t[k] = k < 8 + CNT ? part[0][k] : 0u;
#pragma unroll
  for (int g = 1; g < GROUPS; g++) add_shifted(t, part[g], 8 + CNT, g * CNT);
  return fp_reduce_wide(t);
}
template <int V>
__global__ void k(fp_t *io, int iters, long long *cycles) {
  fp_t x = io[threadIdx.x], y = io[32 + threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (V == 0) x = fp_mul(x, y);
    else if (V == 1) x = fp_mul_ilp<2>(x, y);
    else x = fp_mul_ilp<4>(x, y);
  }
  long long t1 = clock64();
  io[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  fp_t h[64];
  for (int i = 0; i < 64; i++) for (int k = 0; k < 8; k++) h[i].v[k] = 0x9e3779b9u * (i * 8 + k + 1);
  for (int i = 0; i < 64; i++) h[i].v[7] &= 0x7fffffffu;
  fp_t *d; long long *dc;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&dc, 8);
  const int iters = 2000;
  fp_t ref[32];
  for (int v = 0; v < 3; v++) {
    long long best = 1LL << 60;
    fp_t out[32];
    for (int rep = 0; rep < 3; rep++) {
      cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
      if (v == 0) k<0><<<1, 32>>>(d, iters, dc); else if (v == 1) k<1><<<1, 32>>>(d, iters, dc); else k<2><<<1, 32>>>(d, iters, dc);
      long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(out, d, sizeof(out), cudaMemcpyDeviceToHost);
      if (c < best) best = c;
    }
    bool same = true;
    if (v == 0) for (int i = 0; i < 32; i++) ref[i] = out[i];
    else for (int i = 0; i < 32; i++) for (int kk = 0; kk < 8; kk++) same = same && ref[i].v[kk] == out[i].v[kk];
    printf("variant %d: %.1f cycles per dependent F_p multiplication (one warp)  %s\n", v, (double)best / iters, v ? (same ? "results identical" : "RESULTS DIFFER") : "");
  }
  return 0;
}
