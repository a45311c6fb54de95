// F_l table kernels for sm_100a: sumcheck round evaluations, binds, eq tables, CSR/CSC SpMV and the SPARK layer
// builders. HBM-bound integer work: one 32-byte element per 2 x LDG.128, grid-stride loops sized in multiples of the
// 148 SMs, warp-shuffle + shared-memory reductions of field elements. No tensor cores (nothing here is a contraction).
#include "launch_count.hpp"
#include <atomic>

#include "kernels_poly.cuh"

namespace vpin {


__device__ __forceinline__ fl_t ldg_fl(const fl_t *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fl_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ fl_t ld_fl(const fl_t *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1];
  fl_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fl(fl_t *p, const fl_t &x) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ fl_t shfl_down_fl(const fl_t &x, int off) {
  fl_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], off);
  return r;
}

// Block-wide sum of K field elements per thread; thread 0 writes them to dst[0..K).
template <int K>
__device__ __forceinline__ void block_sum_store(fl_t (&acc)[K], fl_t *dst) {
  __shared__ fl_t sm[K][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = fl_add(acc[k], shfl_down_fl(acc[k], off));
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; k++) sm[k][warp] = acc[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      fl_t v = lane < nwarps ? sm[k][lane] : fl_zero();
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v = fl_add(v, shfl_down_fl(v, off));
      if (lane == 0) st_fl(dst + k, v);
    }
  }
}

// second stage: out[inst*K + k] = sum_b partials[(inst*nblocks + b)*K + k]
template <int K>
__global__ void __launch_bounds__(kRedThreads) k_reduce_partials(const fl_t *partials, int nblocks, fl_t *out) {
  const fl_t *p = partials + (size_t)blockIdx.x * nblocks * K;
  fl_t acc[K];
#pragma unroll
  for (int k = 0; k < K; k++) acc[k] = fl_zero();
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = fl_add(acc[k], ld_fl(p + (size_t)b * K + k));
  block_sum_store<K>(acc, out + (size_t)blockIdx.x * K);
}

static inline int red_blocks(size_t n) {
  size_t b = (n + kRedThreads - 1) / kRedThreads;
  if (b < 1) b = 1;
  return (int)(b > (size_t)kRedBlocks ? kRedBlocks : b);
}
static inline unsigned ew_blocks(size_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ eq tables
__global__ void __launch_bounds__(1024) k_eq_small(const fl_t *r, int ell, fl_t *dst, fl_t *tmp) {
  fl_t *a = (ell & 1) ? tmp : dst, *b = (ell & 1) ? dst : tmp;  // ell swaps end in dst
  if (threadIdx.x == 0) st_fl(a, fl_one());
  __syncthreads();
  for (int j = 0; j < ell; j++) {
    fl_t rj = ld_fl(r + j);
    int size = 1 << j;
    for (int i = threadIdx.x; i < size; i += blockDim.x) {
      fl_t s = ld_fl(a + i);
      fl_t hi = fl_mul(s, rj);
      st_fl(b + 2 * i + 1, hi);
      st_fl(b + 2 * i, fl_sub(s, hi));
    }
    __syncthreads();
    fl_t *t = a; a = b; b = t;
  }
}
__global__ void __launch_bounds__(256) k_eq_outer(const fl_t *hi, const fl_t *lo, int lo_bits, size_t n, fl_t *out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_fl(out + i, fl_mul(ldg_fl(hi + (i >> lo_bits)), ldg_fl(lo + (i & (((size_t)1 << lo_bits) - 1)))));
}
// the point in kernel-parameter space (Spartan/src/dense_mlpoly.rs:78-94)
__global__ void __launch_bounds__(1024) k_eq_small_pt(EqPoint pt, int first, int ell, fl_t *dst, fl_t *tmp) {
  fl_t *a = (ell & 1) ? tmp : dst, *b = (ell & 1) ? dst : tmp;
  if (threadIdx.x == 0) st_fl(a, fl_one());
  __syncthreads();
  for (int j = 0; j < ell; j++) {
    fl_t rj = pt.r[first + j];
    int size = 1 << j;
    for (int i = threadIdx.x; i < size; i += blockDim.x) {
      fl_t s = ld_fl(a + i);
      fl_t hi = fl_mul(s, rj);
      st_fl(b + 2 * i + 1, hi);
      st_fl(b + 2 * i, fl_sub(s, hi));
    }
    __syncthreads();
    fl_t *t = a; a = b; b = t;
  }
}
void launch_eq_evals_pt(const EqPoint &pt, int ell, fl_t *d_out, fl_t *d_tmp, cudaStream_t st) {
  if (ell <= 12) {
    ++g_kernel_launches, k_eq_small_pt<<<1, ell <= 8 ? 256 : 1024, 0, st>>>(pt, 0, ell, d_out, d_tmp);
    return;
  }
  int hi_bits = ell / 2, lo_bits = ell - hi_bits;
  fl_t *hi = d_tmp, *lo = d_tmp + ((size_t)1 << hi_bits), *scratch = lo + ((size_t)1 << lo_bits);
  ++g_kernel_launches, k_eq_small_pt<<<1, 1024, 0, st>>>(pt, 0, hi_bits, hi, scratch);
  ++g_kernel_launches, k_eq_small_pt<<<1, 1024, 0, st>>>(pt, hi_bits, lo_bits, lo, scratch);
  size_t n = (size_t)1 << ell;
  unsigned blocks = (unsigned)((n / 256) < 148 * 16 ? (n / 256) : 148 * 16);
  ++g_kernel_launches, k_eq_outer<<<blocks, 256, 0, st>>>(hi, lo, lo_bits, n, d_out);
}
// ---- suffix eq tables: S[2^k + x] = eq(r[ell-k .. ell), x) for k = 0..kmax, x < 2^k (S[1] = 1, S[0] unused).
// The sumcheck kernels factor the shared eq polynomial out of the round polynomial (kernels_round.cu): round j needs
// eq(r[j+1..ell), .), i.e. exactly the tables this doubling construction passes through when the new variable goes on TOP
// of the index, so one pass builds the table of every round (2^kmax+1 elements in all, one multiplication per element).
static const int kEqSuffixSmall = 11;  // levels built by one block
__global__ void __launch_bounds__(1024) k_eq_suffix_small(EqPoint pt, int ell, int kmax, fl_t *S) {
  if (threadIdx.x == 0) st_fl(S + 1, fl_one());
  __syncthreads();
  for (int k = 0; k < kmax; k++) {  // T_{k+1} from T_k: the new top variable is r[ell - 1 - k]
    fl_t w = pt.r[ell - 1 - k];
    int size = 1 << k;
    for (int x = threadIdx.x; x < size; x += blockDim.x) {
      fl_t s = ld_fl(S + size + x);
      fl_t hi = fl_mul(s, w);
      st_fl(S + 2 * size + size + x, hi);
      st_fl(S + 2 * size + x, fl_sub(s, hi));
    }
    __syncthreads();
  }
}
// thread x < 2^k0 expands T_k0[x] into its descendants of the next `levels` (<= 3) tables
__global__ void __launch_bounds__(256) k_eq_suffix_expand(EqPoint pt, int ell, int k0, int levels, fl_t *S) {
  size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n0 = (size_t)1 << k0;
  if (x >= n0) return;
  fl_t v[8];
  v[0] = ld_fl(S + n0 + x);
#pragma unroll
  for (int l = 0; l < 3; l++) {
    if (l >= levels) break;
    fl_t w = pt.r[ell - 1 - (k0 + l)];
    fl_t *T = S + (n0 << (l + 1));  // table k0 + l + 1
#pragma unroll
    for (int y = 0; y < (1 << l); y++) {
      fl_t hi = fl_mul(v[y], w);
      fl_t lo = fl_sub(v[y], hi);
      v[y] = lo;
      v[y + (1 << l)] = hi;
      st_fl(T + x + ((size_t)y << k0), lo);
      st_fl(T + x + ((size_t)(y + (1 << l)) << k0), hi);
    }
  }
}
void launch_eq_suffix(const EqPoint &pt, int ell, int kmax, fl_t *S, cudaStream_t st) {
  int small = kmax < kEqSuffixSmall ? kmax : kEqSuffixSmall;
  ++g_kernel_launches, k_eq_suffix_small<<<1, small <= 8 ? 256 : 1024, 0, st>>>(pt, ell, small, S);
  for (int k0 = small; k0 < kmax; k0 += 3) {
    int levels = kmax - k0 < 3 ? kmax - k0 : 3;
    size_t n0 = (size_t)1 << k0;
    ++g_kernel_launches, k_eq_suffix_expand<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(pt, ell, k0, levels, S);
  }
}
void launch_eq_evals(const fl_t *d_r, int ell, fl_t *d_out, fl_t *d_tmp, cudaStream_t st) {
  if (ell <= 12) {
    ++g_kernel_launches, k_eq_small<<<1, 1024, 0, st>>>(d_r, ell, d_out, d_tmp);
    return;
  }
  int hi_bits = ell / 2, lo_bits = ell - hi_bits;
  fl_t *hi = d_tmp, *lo = d_tmp + ((size_t)1 << hi_bits), *scratch = lo + ((size_t)1 << lo_bits);
  ++g_kernel_launches, k_eq_small<<<1, 1024, 0, st>>>(d_r, hi_bits, hi, scratch);
  ++g_kernel_launches, k_eq_small<<<1, 1024, 0, st>>>(d_r + hi_bits, lo_bits, lo, scratch);
  size_t n = (size_t)1 << ell;
  unsigned blocks = (unsigned)((n / 256) < 148 * 16 ? (n / 256) : 148 * 16);
  ++g_kernel_launches, k_eq_outer<<<blocks, 256, 0, st>>>(hi, lo, lo_bits, n, d_out);
}

// ------------------------------------------------------------------------------------------------ binds
__global__ void __launch_bounds__(256) k_bind_top(fl_t *Z, size_t half, const fl_t *d_r) {
  fl_t r = ld_fl(d_r);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    fl_t lo = ld_fl(Z + i), hi = ld_fl(Z + half + i);
    st_fl(Z + i, fl_add(lo, fl_mul(r, fl_sub(hi, lo))));
  }
}
__global__ void __launch_bounds__(256) k_bind_top_multi(fl_t *const *tables, size_t half, const fl_t *d_r) {
  fl_t r = ld_fl(d_r);
  fl_t *Z = tables[blockIdx.y];
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    fl_t lo = ld_fl(Z + i), hi = ld_fl(Z + half + i);
    st_fl(Z + i, fl_add(lo, fl_mul(r, fl_sub(hi, lo))));
  }
}
__global__ void __launch_bounds__(256) k_bind_bot(const fl_t *Z, fl_t *out, size_t half, const fl_t *d_r) {
  fl_t r = ld_fl(d_r);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    fl_t lo = ld_fl(Z + 2 * i), hi = ld_fl(Z + 2 * i + 1);
    st_fl(out + i, fl_add(lo, fl_mul(r, fl_sub(hi, lo))));
  }
}
static inline unsigned stream_blocks(size_t n) {
  size_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  size_t cap = 148 * 8;
  return (unsigned)(b > cap ? cap : b);
}
void launch_bind_top(fl_t *Z, size_t half, const fl_t *d_r, cudaStream_t st) {
  ++g_kernel_launches, k_bind_top<<<stream_blocks(half), 256, 0, st>>>(Z, half, d_r);
}
void launch_bind_top_multi(fl_t *const *d_tables, int ntables, size_t half, const fl_t *d_r, cudaStream_t st) {
  dim3 grid(stream_blocks(half), ntables);
  ++g_kernel_launches, k_bind_top_multi<<<grid, 256, 0, st>>>(d_tables, half, d_r);
}
void launch_bind_bot(const fl_t *Z, fl_t *out, size_t half, const fl_t *d_r, cudaStream_t st) {
  ++g_kernel_launches, k_bind_bot<<<stream_blocks(half), 256, 0, st>>>(Z, out, half, d_r);
}

// ------------------------------------------------------------------------------------------------ sumcheck rounds
__global__ void __launch_bounds__(kRedThreads) k_cubic_additive(const fl_t *A, const fl_t *B, const fl_t *C, const fl_t *D,
                                                                size_t half, fl_t *partials) {
  fl_t acc[3] = {fl_zero(), fl_zero(), fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    fl_t a0 = ldg_fl(A + i), a1 = ldg_fl(A + half + i);
    fl_t b0 = ldg_fl(B + i), b1 = ldg_fl(B + half + i);
    fl_t c0 = ldg_fl(C + i), c1 = ldg_fl(C + half + i);
    fl_t d0 = ldg_fl(D + i), d1 = ldg_fl(D + half + i);
    acc[0] = fl_add(acc[0], fl_mul(a0, fl_sub(fl_mul(b0, c0), d0)));
    fl_t da = fl_sub(a1, a0), db = fl_sub(b1, b0), dc = fl_sub(c1, c0), dd = fl_sub(d1, d0);
    fl_t a2 = fl_add(a1, da), b2 = fl_add(b1, db), c2 = fl_add(c1, dc), d2 = fl_add(d1, dd);
    acc[1] = fl_add(acc[1], fl_mul(a2, fl_sub(fl_mul(b2, c2), d2)));
    fl_t a3 = fl_add(a2, da), b3 = fl_add(b2, db), c3 = fl_add(c2, dc), d3 = fl_add(d2, dd);
    acc[2] = fl_add(acc[2], fl_mul(a3, fl_sub(fl_mul(b3, c3), d3)));
  }
  block_sum_store<3>(acc, partials + (size_t)blockIdx.x * 3);
}
void launch_cubic_additive_round(const fl_t *A, const fl_t *B, const fl_t *C, const fl_t *D, size_t half, fl_t *d_out,
                                 fl_t *d_partials, cudaStream_t st) {
  int nb = red_blocks(half);
  ++g_kernel_launches, k_cubic_additive<<<nb, kRedThreads, 0, st>>>(A, B, C, D, half, d_partials);
  ++g_kernel_launches, k_reduce_partials<3><<<1, kRedThreads, 0, st>>>(d_partials, nb, d_out);
}

__global__ void __launch_bounds__(kRedThreads) k_quad(const fl_t *A, const fl_t *B, size_t half, fl_t *partials) {
  fl_t acc[2] = {fl_zero(), fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    fl_t a0 = ldg_fl(A + i), a1 = ldg_fl(A + half + i);
    fl_t b0 = ldg_fl(B + i), b1 = ldg_fl(B + half + i);
    acc[0] = fl_add(acc[0], fl_mul(a0, b0));
    fl_t a2 = fl_add(a1, fl_sub(a1, a0)), b2 = fl_add(b1, fl_sub(b1, b0));
    acc[1] = fl_add(acc[1], fl_mul(a2, b2));
  }
  block_sum_store<2>(acc, partials + (size_t)blockIdx.x * 2);
}
void launch_quad_round(const fl_t *A, const fl_t *B, size_t half, fl_t *d_out, fl_t *d_partials, cudaStream_t st) {
  int nb = red_blocks(half);
  ++g_kernel_launches, k_quad<<<nb, kRedThreads, 0, st>>>(A, B, half, d_partials);
  ++g_kernel_launches, k_reduce_partials<2><<<1, kRedThreads, 0, st>>>(d_partials, nb, d_out);
}

__global__ void __launch_bounds__(kRedThreads) k_cubic_batched(const fl_t *const *pA, const fl_t *const *pB, const fl_t *const *pC,
                                                               size_t half, fl_t *partials) {
  const fl_t *A = pA[blockIdx.y], *B = pB[blockIdx.y], *C = pC[blockIdx.y];
  fl_t acc[3] = {fl_zero(), fl_zero(), fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    fl_t a0 = ldg_fl(A + i), a1 = ldg_fl(A + half + i);
    fl_t b0 = ldg_fl(B + i), b1 = ldg_fl(B + half + i);
    fl_t c0 = ldg_fl(C + i), c1 = ldg_fl(C + half + i);
    acc[0] = fl_add(acc[0], fl_mul(fl_mul(a0, b0), c0));
    fl_t da = fl_sub(a1, a0), db = fl_sub(b1, b0), dc = fl_sub(c1, c0);
    fl_t a2 = fl_add(a1, da), b2 = fl_add(b1, db), c2 = fl_add(c1, dc);
    acc[1] = fl_add(acc[1], fl_mul(fl_mul(a2, b2), c2));
    fl_t a3 = fl_add(a2, da), b3 = fl_add(b2, db), c3 = fl_add(c2, dc);
    acc[2] = fl_add(acc[2], fl_mul(fl_mul(a3, b3), c3));
  }
  block_sum_store<3>(acc, partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3);
}
void launch_cubic_batched_round(const fl_t *const *d_A, const fl_t *const *d_B, const fl_t *const *d_C, int n, size_t half,
                                fl_t *d_out, fl_t *d_partials, cudaStream_t st) {
  int nb = red_blocks(half);
  if (nb > kRedBlocks / 4 && n > 4) nb = kRedBlocks / 4;  // n instances share the machine
  dim3 grid(nb, n);
  ++g_kernel_launches, k_cubic_batched<<<grid, kRedThreads, 0, st>>>(d_A, d_B, d_C, half, d_partials);
  ++g_kernel_launches, k_reduce_partials<3><<<n, kRedThreads, 0, st>>>(d_partials, nb, d_out);
}

// ------------------------------------------------------------------------------------------------ dot products
__global__ void __launch_bounds__(kRedThreads) k_dot_multi(const fl_t *const *pA, const fl_t *A0, const fl_t *B, size_t n,
                                                           fl_t *partials) {
  const fl_t *A = pA ? pA[blockIdx.y] : A0;
  fl_t acc[1] = {fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc[0] = fl_add(acc[0], fl_mul(ldg_fl(A + i), ldg_fl(B + i)));
  block_sum_store<1>(acc, partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x));
}
void launch_dot(const fl_t *A, const fl_t *B, size_t n, fl_t *d_out, fl_t *d_partials, cudaStream_t st) {
  int nb = red_blocks(n);
  ++g_kernel_launches, k_dot_multi<<<dim3(nb, 1), kRedThreads, 0, st>>>(nullptr, A, B, n, d_partials);
  ++g_kernel_launches, k_reduce_partials<1><<<1, kRedThreads, 0, st>>>(d_partials, nb, d_out);
}
void launch_dot_multi(const fl_t *const *d_A, const fl_t *B, int n, size_t len, fl_t *d_out, fl_t *d_partials, cudaStream_t st) {
  int nb = red_blocks(len);
  if (nb > kRedBlocks / 4 && n > 4) nb = kRedBlocks / 4;
  ++g_kernel_launches, k_dot_multi<<<dim3(nb, n), kRedThreads, 0, st>>>(d_A, nullptr, B, len, d_partials);
  ++g_kernel_launches, k_reduce_partials<1><<<n, kRedThreads, 0, st>>>(d_partials, nb, d_out);
}
__global__ void __launch_bounds__(kRedThreads) k_dot3(const fl_t *A, const fl_t *B, const fl_t *C, size_t n, fl_t *partials) {
  fl_t acc[1] = {fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc[0] = fl_add(acc[0], fl_mul(fl_mul(ldg_fl(A + i), ldg_fl(B + i)), ldg_fl(C + i)));
  block_sum_store<1>(acc, partials + blockIdx.x);
}
void launch_dot3(const fl_t *A, const fl_t *B, const fl_t *C, size_t n, fl_t *d_out, fl_t *d_partials, cudaStream_t st) {
  int nb = red_blocks(n);
  ++g_kernel_launches, k_dot3<<<nb, kRedThreads, 0, st>>>(A, B, C, n, d_partials);
  ++g_kernel_launches, k_reduce_partials<1><<<1, kRedThreads, 0, st>>>(d_partials, nb, d_out);
}

// ------------------------------------------------------------------------------------------------ L * Z
// grid (ceil(R/128), nsplit): thread owns column j, walks its slice of the L rows; slices summed by k_bound_finish
__global__ void __launch_bounds__(128) k_bound(const fl_t *Z, const fl_t *L, size_t Lsize, size_t Rsize, fl_t *tmp) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Rsize) return;
  size_t per = (Lsize + gridDim.y - 1) / gridDim.y;
  size_t lo = per * blockIdx.y, hi = lo + per < Lsize ? lo + per : Lsize;
  fl_t acc = fl_zero();
  size_t i = lo;
  for (; i + 4 <= hi; i += 4) {  // four rows' loads in flight before the first multiplication (the chain through acc hid none of them)
    fl_t z0 = ldg_fl(Z + i * Rsize + j), z1 = ldg_fl(Z + (i + 1) * Rsize + j), z2 = ldg_fl(Z + (i + 2) * Rsize + j),
         z3 = ldg_fl(Z + (i + 3) * Rsize + j);
    fl_t l0 = ldg_fl(L + i), l1 = ldg_fl(L + i + 1), l2 = ldg_fl(L + i + 2), l3 = ldg_fl(L + i + 3);
    acc = fl_add(acc, fl_add(fl_add(fl_mul(l0, z0), fl_mul(l1, z1)), fl_add(fl_mul(l2, z2), fl_mul(l3, z3))));
  }
  for (; i < hi; i++) acc = fl_add(acc, fl_mul(ldg_fl(L + i), ldg_fl(Z + i * Rsize + j)));
  st_fl(tmp + (size_t)blockIdx.y * Rsize + j, acc);
}
__global__ void __launch_bounds__(128) k_bound_finish(const fl_t *tmp, int nsplit, size_t Rsize, fl_t *out) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Rsize) return;
  fl_t acc = fl_zero();
  for (int s = 0; s < nsplit; s++) acc = fl_add(acc, ld_fl(tmp + (size_t)s * Rsize + j));
  st_fl(out + j, acc);
}
void launch_bound(const fl_t *Z, const fl_t *L, size_t Lsize, size_t Rsize, fl_t *d_out, fl_t *d_tmp, cudaStream_t st) {
  unsigned bx = (unsigned)((Rsize + 127) / 128);
  int nsplit = 1;
  while (nsplit < 64 && (size_t)bx * nsplit < 148 * 8 && (size_t)nsplit * 2 <= Lsize) nsplit *= 2;
  ++g_kernel_launches, k_bound<<<dim3(bx, nsplit), 128, 0, st>>>(Z, L, Lsize, Rsize, d_tmp);
  ++g_kernel_launches, k_bound_finish<<<bx, 128, 0, st>>>(d_tmp, nsplit, Rsize, d_out);
}

// ------------------------------------------------------------------------------------------------ SpMV
// acc += coefficient * x with the coefficient given by its dictionary code; the 32-byte value is read only for kCodeGeneral
__device__ __forceinline__ fl_t spmv_term(const fl_t &acc, uint8_t code, const fl_t *val, uint32_t e, const fl_t &x) {
  switch (code) {
    case kCodePlus1: return fl_add(acc, x);
    case kCodeMinus1: return fl_sub(acc, x);
    case kCodePlus2: return fl_add(acc, fl_dbl(x));
    case kCodeMinus2: return fl_sub(acc, fl_dbl(x));
    case kCodePlus3: return fl_add(acc, fl_add(fl_dbl(x), x));
    default: return fl_add(acc, fl_mul(ldg_fl(val + e), x));
  }
}
__global__ void __launch_bounds__(256) k_spmv_csr(CsrDev m, const fl_t *z, fl_t *out, int skip_long) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m.n) return;
  uint32_t e = m.ptr[row], end = m.ptr[row + 1];
  if (skip_long && end - e > (uint32_t)kLongRow) return;  // done by k_spmv_csr_long
  fl_t acc = fl_zero();
  for (; e < end; e++) acc = spmv_term(acc, m.code[e], m.val, e, ldg_fl(z + m.idx[e]));
  st_fl(out + row, acc);
}
// one warp per long row: lanes stride the entries (coalesced code / index / value loads), shuffle-tree sum
__global__ void __launch_bounds__(128) k_spmv_csr_long(CsrDev m, const fl_t *z, fl_t *out) {
  size_t w = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= m.n_long_rows) return;
  uint32_t row = m.long_rows[w];
  fl_t acc = fl_zero();
  for (uint32_t e = m.ptr[row] + lane, end = m.ptr[row + 1]; e < end; e += 32) acc = spmv_term(acc, m.code[e], m.val, e, ldg_fl(z + m.idx[e]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    fl_t o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.v[i] = __shfl_down_sync(0xffffffffu, acc.v[i], off);
    acc = fl_add(acc, o);
  }
  if (lane == 0) st_fl(out + row, acc);
}
void launch_spmv_csr(const CsrDev &m, const fl_t *z, fl_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_spmv_csr<<<ew_blocks(m.n), 256, 0, st>>>(m, z, out, m.n_long_rows ? 1 : 0);
  if (m.n_long_rows) ++g_kernel_launches, k_spmv_csr_long<<<(unsigned)((m.n_long_rows + 3) / 4), 128, 0, st>>>(m, z, out);
}
__global__ void __launch_bounds__(256) k_spmv_csc(CscDev m, const fl_t *x, const fl_t *d_scale, int accumulate, fl_t *out) {
  size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= m.n) return;
  uint32_t e = m.ptr[col], end = m.ptr[col + 1];
  if (end - e > (uint32_t)kLongCol) return;  // done by k_spmv_csc_long
  fl_t acc = fl_zero();
  for (; e < end; e++) acc = spmv_term(acc, m.code[e], m.val, e, ldg_fl(x + m.idx[e]));
  acc = fl_mul(acc, ld_fl(d_scale));
  if (accumulate) acc = fl_add(acc, ld_fl(out + col));
  st_fl(out + col, acc);
}
// long columns (the constant-1 column of vPIN's instances holds ~15 entries per bit-step of every multiplication): block
// (c, s) adds slice s of column long_cols[c]; k_spmv_csc_long_fin adds the slices, scales and stores
__global__ void __launch_bounds__(kRedThreads) k_spmv_csc_long(CscDev m, const fl_t *x, fl_t *scratch) {
  uint32_t col = m.long_cols[blockIdx.x];
  uint32_t beg = m.ptr[col], end = m.ptr[col + 1];
  uint32_t per = (end - beg + gridDim.y - 1) / gridDim.y;
  uint32_t lo = beg + per * blockIdx.y, hi = lo + per < end ? lo + per : end;
  fl_t acc[1] = {fl_zero()};
  for (uint32_t e = lo + threadIdx.x; e < hi; e += blockDim.x) acc[0] = spmv_term(acc[0], m.code[e], m.val, e, ldg_fl(x + m.idx[e]));
  block_sum_store<1>(acc, scratch + (size_t)blockIdx.x * gridDim.y + blockIdx.y);
}
__global__ void __launch_bounds__(64) k_spmv_csc_long_fin(CscDev m, const fl_t *scratch, int nsplit, const fl_t *d_scale, int accumulate,
                                                          fl_t *out) {
  uint32_t col = m.long_cols[blockIdx.x];
  fl_t acc[1] = {fl_zero()};
  for (int s = threadIdx.x; s < nsplit; s += blockDim.x) acc[0] = fl_add(acc[0], ld_fl(scratch + (size_t)blockIdx.x * nsplit + s));
  __shared__ fl_t res;
  block_sum_store<1>(acc, &res);
  __syncthreads();
  if (threadIdx.x == 0) {
    fl_t v = fl_mul(res, ld_fl(d_scale));
    if (accumulate) v = fl_add(v, ld_fl(out + col));
    st_fl(out + col, v);
  }
}
void launch_spmv_csc_scaled(const CscDev &m, const fl_t *x, const fl_t *d_scale, bool accumulate, fl_t *out, fl_t *d_scratch,
                            size_t scratch_elems, cudaStream_t st) {
  ++g_kernel_launches, k_spmv_csc<<<ew_blocks(m.n), 256, 0, st>>>(m, x, d_scale, accumulate ? 1 : 0, out);
  if (m.n_long) {
    size_t nsplit = scratch_elems / m.n_long;
    if (nsplit > 64) nsplit = 64;
    if (nsplit < 1) nsplit = 1;  // the caller sizes the scratch for at least one slot per long column
    ++g_kernel_launches, k_spmv_csc_long<<<dim3((unsigned)m.n_long, (unsigned)nsplit), kRedThreads, 0, st>>>(m, x, d_scratch);
    ++g_kernel_launches, k_spmv_csc_long_fin<<<(unsigned)m.n_long, 64, 0, st>>>(m, d_scratch, (int)nsplit, d_scale, accumulate ? 1 : 0, out);
  }
}
__global__ void __launch_bounds__(kRedThreads) k_sparse_eval(const uint32_t *rows, const uint32_t *cols, const fl_t *val, size_t nnz,
                                                             const fl_t *trx, const fl_t *try_, fl_t *partials) {
  fl_t acc[1] = {fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride)
    acc[0] = fl_add(acc[0], fl_mul(fl_mul(ldg_fl(trx + rows[i]), ldg_fl(try_ + cols[i])), ldg_fl(val + i)));
  block_sum_store<1>(acc, partials + blockIdx.x);
}
void launch_sparse_eval(const uint32_t *rows, const uint32_t *cols, const fl_t *val, size_t nnz, const fl_t *trx, const fl_t *try_,
                        fl_t *d_out, fl_t *d_partials, cudaStream_t st) {
  int nb = red_blocks(nnz);
  ++g_kernel_launches, k_sparse_eval<<<nb, kRedThreads, 0, st>>>(rows, cols, val, nnz, trx, try_, d_partials);
  ++g_kernel_launches, k_reduce_partials<1><<<1, kRedThreads, 0, st>>>(d_partials, nb, d_out);
}

// ------------------------------------------------------------------------------------------------ SPARK layers
__global__ void __launch_bounds__(256) k_gather(const uint32_t *addr, const fl_t *mem, size_t n, fl_t *out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_fl(out + i, ldg_fl(mem + addr[i]));
}
void launch_gather(const uint32_t *addr, const fl_t *mem, size_t n, fl_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_gather<<<stream_blocks(n), 256, 0, st>>>(addr, mem, n, out);
}
// Montgomery form of a 32-bit integer: x * R mod l == mont_mul(x, R^2)
__device__ __forceinline__ fl_t fl_from_u32_dev(uint32_t x) {
  fl_t t = fl_zero();
  t.v[0] = x;
  return fl_mul(t, fl_r2());
}
__global__ void __launch_bounds__(256) k_u32_to_fl(const uint32_t *in, size_t n, fl_t *out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_fl(out + i, fl_from_u32_dev(in[i]));
}
void launch_u32_to_fl(const uint32_t *in, size_t n, fl_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_u32_to_fl<<<stream_blocks(n), 256, 0, st>>>(in, n, out);
}
__global__ void __launch_bounds__(256) k_hash_mem(const fl_t *eq, const uint32_t *audit_ts, size_t n, const fl_t *d_gt, fl_t *init,
                                                  fl_t *audit) {
  fl_t gamma = ld_fl(d_gt), tau = ld_fl(d_gt + 1);
  fl_t gamma2 = fl_mul(gamma, gamma);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // hash(addr, val, ts) = ts*gamma^2 + val*gamma + addr  (sparse_mlpoly.rs:562-566)
    fl_t base = fl_sub(fl_add(fl_mul(ldg_fl(eq + i), gamma), fl_from_u32_dev((uint32_t)i)), tau);
    st_fl(init + i, base);
    st_fl(audit + i, fl_add(fl_mul(fl_from_u32_dev(audit_ts[i]), gamma2), base));
  }
}
void launch_hash_mem(const fl_t *eq, const uint32_t *audit_ts, size_t num_cells, const fl_t *d_gt, fl_t *init, fl_t *audit, cudaStream_t st) {
  ++g_kernel_launches, k_hash_mem<<<stream_blocks(num_cells), 256, 0, st>>>(eq, audit_ts, num_cells, d_gt, init, audit);
}
__global__ void __launch_bounds__(256) k_hash_ops(const uint32_t *addr, const fl_t *deref, const uint32_t *read_ts, size_t n,
                                                  const fl_t *d_gt, fl_t *read, fl_t *write) {
  fl_t gamma = ld_fl(d_gt), tau = ld_fl(d_gt + 1);
  fl_t gamma2 = fl_mul(gamma, gamma);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    fl_t base = fl_sub(fl_add(fl_mul(ldg_fl(deref + i), gamma), fl_from_u32_dev(addr[i])), tau);
    fl_t r = fl_add(fl_mul(fl_from_u32_dev(read_ts[i]), gamma2), base);
    st_fl(read + i, r);
    st_fl(write + i, fl_add(r, gamma2));  // write_ts = read_ts + 1
  }
}
void launch_hash_ops(const uint32_t *addr, const fl_t *deref, const uint32_t *read_ts, size_t num_ops, const fl_t *d_gt, fl_t *read,
                     fl_t *write, cudaStream_t st) {
  ++g_kernel_launches, k_hash_ops<<<stream_blocks(num_ops), 256, 0, st>>>(addr, deref, read_ts, num_ops, d_gt, read, write);
}
__global__ void __launch_bounds__(256) k_mul_halves(const fl_t *in, size_t n, fl_t *out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_fl(out + i, fl_mul(ldg_fl(in + i), ldg_fl(in + n + i)));
}
void launch_mul_halves(const fl_t *in, size_t n, fl_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_mul_halves<<<stream_blocks(n), 256, 0, st>>>(in, n, out);
}
// the same layer of several packed trees in one launch (tree = blockIdx.y)
__global__ void __launch_bounds__(256) k_mul_halves_multi(TreeBatch b, size_t off, size_t half) {
  fl_t *t = b.p[blockIdx.y];
  const fl_t *in = t + off;
  fl_t *out = t + off + 2 * half;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride)
    st_fl(out + i, fl_mul(ldg_fl(in + i), ldg_fl(in + half + i)));
}
// every remaining layer of a tree (current layer of `vlen` <= kTreeTail elements at `off`) by ONE block per tree: the upper
// layers are a chain of ever shorter launches otherwise (12 trees x 13 layers of <= 4096 products each)
__global__ void __launch_bounds__(1024) k_tree_tail(TreeBatch b, size_t off, size_t vlen) {
  fl_t *t = b.p[blockIdx.x];
  while (vlen > 2) {
    size_t half = vlen / 2;
    for (size_t i = threadIdx.x; i < half; i += blockDim.x) {
      const uint4 *pa = reinterpret_cast<const uint4 *>(t + off + i), *pb = reinterpret_cast<const uint4 *>(t + off + half + i);
      fl_t x, y;
      uint4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];  // plain loads: the layer below was written by this block
      x.v[0] = a0.x; x.v[1] = a0.y; x.v[2] = a0.z; x.v[3] = a0.w; x.v[4] = a1.x; x.v[5] = a1.y; x.v[6] = a1.z; x.v[7] = a1.w;
      y.v[0] = b0.x; y.v[1] = b0.y; y.v[2] = b0.z; y.v[3] = b0.w; y.v[4] = b1.x; y.v[5] = b1.y; y.v[6] = b1.z; y.v[7] = b1.w;
      st_fl(t + off + vlen + i, fl_mul(x, y));
    }
    __syncthreads();
    off += vlen;
    vlen = half;
  }
}
void launch_build_trees(const TreeBatch &b, size_t n, cudaStream_t st) {
  size_t off = 0, vlen = n;
  for (; vlen > kTreeTail; vlen /= 2) {
    dim3 grid(stream_blocks(vlen / 2), b.n);
    ++g_kernel_launches, k_mul_halves_multi<<<grid, 256, 0, st>>>(b, off, vlen / 2);
    off += vlen;
  }
  if (vlen > 2) ++g_kernel_launches, k_tree_tail<<<b.n, 1024, 0, st>>>(b, off, vlen);
}
__global__ void __launch_bounds__(256) k_lincomb3(const fl_t *A, const fl_t *B, const fl_t *C, const fl_t *d_abc, size_t n, fl_t *out) {
  fl_t a = ld_fl(d_abc), b = ld_fl(d_abc + 1), c = ld_fl(d_abc + 2);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_fl(out + i, fl_add(fl_add(fl_mul(a, ldg_fl(A + i)), fl_mul(b, ldg_fl(B + i))), fl_mul(c, ldg_fl(C + i))));
}
void launch_lincomb3(const fl_t *A, const fl_t *B, const fl_t *C, const fl_t *d_abc, size_t n, fl_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_lincomb3<<<stream_blocks(n), 256, 0, st>>>(A, B, C, d_abc, n, out);
}

// ------------------------------------------------------------------------------------------------ bullet reduction
__global__ void __launch_bounds__(256) k_bullet_fold(fl_t *a, fl_t *b, size_t n, const fl_t *d_u) {
  fl_t u = ld_fl(d_u), uinv = ld_fl(d_u + 1);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    st_fl(a + i, fl_add(fl_mul(ld_fl(a + i), u), fl_mul(uinv, ld_fl(a + n + i))));
    st_fl(b + i, fl_add(fl_mul(ld_fl(b + i), uinv), fl_mul(u, ld_fl(b + n + i))));
  }
}
void launch_bullet_fold(fl_t *a, fl_t *b, size_t n, const fl_t *d_u, cudaStream_t st) {
  ++g_kernel_launches, k_bullet_fold<<<stream_blocks(n), 256, 0, st>>>(a, b, n, d_u);
}
__global__ void __launch_bounds__(256) k_bullet_weights(fl_t *w, size_t total, size_t n, const fl_t *d_u) {
  fl_t u = ld_fl(d_u), uinv = ld_fl(d_u + 1);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride)
    st_fl(w + j, fl_mul(ld_fl(w + j), (j & (2 * n - 1)) < n ? uinv : u));
}
void launch_bullet_weights(fl_t *w, size_t total, size_t n, const fl_t *d_u, cudaStream_t st) {
  ++g_kernel_launches, k_bullet_weights<<<stream_blocks(total), 256, 0, st>>>(w, total, n, d_u);
}
__global__ void __launch_bounds__(256) k_bullet_scalars(const fl_t *a, const fl_t *w, size_t total, size_t n, fl_t *sL, fl_t *sR) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
    size_t rem = j & (2 * n - 1);
    fl_t wj = ld_fl(w + j);
    if (rem >= n) {
      st_fl(sL + j, fl_mul(ld_fl(a + rem - n), wj));  // a_L against G_R
      st_fl(sR + j, fl_zero());
    } else {
      st_fl(sL + j, fl_zero());
      st_fl(sR + j, fl_mul(ld_fl(a + rem + n), wj));  // a_R against G_L
    }
  }
}
void launch_bullet_scalars(const fl_t *a, const fl_t *w, size_t total, size_t n, fl_t *sL, fl_t *sR, cudaStream_t st) {
  ++g_kernel_launches, k_bullet_scalars<<<stream_blocks(total), 256, 0, st>>>(a, w, total, n, sL, sR);
}
__global__ void __launch_bounds__(256) k_scale(const fl_t *in, const fl_t *d_s, size_t n, fl_t *out) {
  fl_t s = ld_fl(d_s);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_fl(out + i, fl_mul(s, ld_fl(in + i)));
}
void launch_scale(const fl_t *in, const fl_t *d_s, size_t n, fl_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_scale<<<stream_blocks(n), 256, 0, st>>>(in, d_s, n, out);
}
__global__ void __launch_bounds__(256) k_add_vec(const fl_t *a, const fl_t *b, size_t n, fl_t *out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_fl(out + i, fl_add(ld_fl(a + i), ld_fl(b + i)));
}
void launch_add_vec(const fl_t *a, const fl_t *b, size_t n, fl_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_add_vec<<<stream_blocks(n), 256, 0, st>>>(a, b, n, out);
}
__global__ void __launch_bounds__(256) k_fill_one(fl_t *out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_fl(out + i, fl_one());
}
void launch_fill_one(fl_t *out, size_t n, cudaStream_t st) { ++g_kernel_launches, k_fill_one<<<stream_blocks(n), 256, 0, st>>>(out, n); }
__global__ void __launch_bounds__(256) k_mont_conv(const fl_t *in, size_t n, fl_t *out, int to_mont) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    fl_t x = ld_fl(in + i);
    st_fl(out + i, to_mont ? fl_to_mont(x) : fl_from_mont(x));
  }
}
// Scalar::from_bytes on the device (Spartan/src/scalar/ristretto255.rs:398-424): canonical limbs -> Montgomery; any value >= l
// raises *bad (the C ABI then reports InvalidScalar)
__global__ void __launch_bounds__(256) k_from_bytes_checked(const fl_t *in, size_t n, fl_t *out, uint32_t *bad) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t any_bad = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    fl_t x = ld_fl(in + i);
    int64_t br = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { int64_t d = (int64_t)x.v[k] - (int64_t)fl_modulus_limb(k) + br; br = d >> 32; }
    any_bad |= br == 0 ? 1u : 0u;  // no borrow -> x >= l
    st_fl(out + i, fl_to_mont(x));
  }
  if (any_bad) atomicOr(bad, 1u);
}
void launch_from_bytes_checked(const fl_t *in, size_t n, fl_t *out, uint32_t *d_bad, cudaStream_t st) {
  ++g_kernel_launches, k_from_bytes_checked<<<stream_blocks(n), 256, 0, st>>>(in, n, out, d_bad);
}
void launch_to_mont(const fl_t *in, size_t n, fl_t *out, cudaStream_t st) { ++g_kernel_launches, k_mont_conv<<<stream_blocks(n), 256, 0, st>>>(in, n, out, 1); }
void launch_from_mont(const fl_t *in, size_t n, fl_t *out, cudaStream_t st) { ++g_kernel_launches, k_mont_conv<<<stream_blocks(n), 256, 0, st>>>(in, n, out, 0); }

}  // namespace vpin
