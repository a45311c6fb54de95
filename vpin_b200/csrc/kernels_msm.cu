// Fixed-base MSM kernels for sm_100a (see kernels_msm.cuh for the design). Integer-multiply bound: the hot loop is
// ge_madd (7 F_p multiplications of 8x8 32-bit limbs) fed by 96-byte table entries fetched with LDG.128.
#include "launch_count.hpp"
#include <atomic>
#include <cstdlib>

#include "fp29.cuh"
#include "kernels_msm.cuh"
#include "msm_recode.cuh"

namespace vpin {


__device__ __forceinline__ fp_t ldg_fp(const fp_t *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fp_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ fp_t ld_fp(const fp_t *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1];
  fp_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fp(fp_t *p, const fp_t &x) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ ge_t ld_ge(const ge_t *p) {
  ge_t r;
  r.X = ld_fp(&p->X); r.Y = ld_fp(&p->Y); r.Z = ld_fp(&p->Z); r.T = ld_fp(&p->T);
  return r;
}
__device__ __forceinline__ void st_ge(ge_t *p, const ge_t &g) { st_fp(&p->X, g.X); st_fp(&p->Y, g.Y); st_fp(&p->Z, g.Z); st_fp(&p->T, g.T); }
__device__ __forceinline__ niels_t ldg_niels(const niels_t *p) {
  niels_t r;
  r.yp = ldg_fp(&p->yp); r.ym = ldg_fp(&p->ym); r.t2d = ldg_fp(&p->t2d);
  return r;
}
__device__ __forceinline__ niels_t ld_niels(const niels_t *p) {
  niels_t r;
  r.yp = ld_fp(&p->yp); r.ym = ld_fp(&p->ym); r.t2d = ld_fp(&p->t2d);
  return r;
}
__device__ __forceinline__ void st_niels(niels_t *p, const niels_t &n) { st_fp(&p->yp, n.yp); st_fp(&p->ym, n.ym); st_fp(&p->t2d, n.t2d); }

// ------------------------------------------------------------------------------------------------ table build
// step 1: table[j][0] = affine Niels form of base j
__global__ void __launch_bounds__(128) k_bases_to_niels(const ge_t *bases, size_t n, int tsize, niels_t *table) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  ge_t g = ld_ge(bases + j);
  st_niels(table + j * tsize, ge_to_niels(g, fp_invert(g.Z)));
}
// step 2: thread (j, c) fills multiples c*CH+1 .. c*CH+CH of base j; one inversion per chunk (Montgomery's trick)
static const int kTblChunk = 32;
__global__ void __launch_bounds__(128) k_table_fill(size_t n, int tsize, niels_t *table) {
  const int chunks = tsize / kTblChunk;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * chunks) return;
  size_t j = tid / chunks;
  int c = (int)(tid % chunks);
  niels_t *slots = table + j * tsize;
  niels_t g1 = ld_niels(slots);
  int first = c == 0 ? 1 : 0;  // slot 0 is already final and is being read by the other chunks
  // P = (c*CH + 1 + first) * G by double-and-add
  uint32_t k = (uint32_t)(c * kTblChunk + 1 + first);
  ge_t P = ge_identity();
  for (int b = 31 - __clz(k); b >= 0; b--) {
    P = ge_dbl(P);
    if ((k >> b) & 1) P = ge_madd(P, g1);
  }
  fp_t prefix[kTblChunk];
  fp_t run = fp_one();
  for (int m = first; m < kTblChunk; m++) {
    niels_t raw;
    raw.yp = P.X; raw.ym = P.Y; raw.t2d = P.Z;
    st_niels(slots + c * kTblChunk + m, raw);
    run = fp_mul(run, P.Z);
    prefix[m] = run;
    P = ge_madd(P, g1);
  }
  fp_t inv = fp_invert(run);
  for (int m = kTblChunk - 1; m >= first; m--) {
    niels_t raw = ld_niels(slots + c * kTblChunk + m);
    fp_t zinv = m > first ? fp_mul(inv, prefix[m - 1]) : inv;
    inv = fp_mul(inv, raw.t2d);
    ge_t q;
    q.X = raw.yp; q.Y = raw.ym; q.Z = raw.t2d; q.T = fp_zero();
    st_niels(slots + c * kTblChunk + m, niels_half(ge_to_niels(q, zinv)));
  }
}
// step 3: slot 0 (read in its full form by every chunk of step 2) is halved last
__global__ void __launch_bounds__(128) k_table_halve_first(size_t n, int tsize, niels_t *table) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  st_niels(table + j * tsize, niels_half(ld_niels(table + j * tsize)));
}
// bases_out[j] = 2^ndbl * bases_in[j]
__global__ void __launch_bounds__(128) k_points_dbl_n(const ge_t *in, size_t n, int ndbl, ge_t *out) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  ge_t g = ld_ge(in + j);
  for (int i = 0; i < ndbl; i++) g = ge_dbl(g);
  st_ge(out + j, g);
}
void launch_table_build(const ge_t *d_bases, size_t n, const MsmGeom &g, niels_t *d_table, ge_t *d_scratch, cudaStream_t st) {
  const ge_t *cur = d_bases;
  size_t threads = n * (g.table / kTblChunk);
  for (int t = 0; t < g.sub; t++) {
    if (t > 0) {
      ++g_kernel_launches, k_points_dbl_n<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cur, n, g.W * g.group, d_scratch);
      cur = d_scratch;
    }
    niels_t *tbl = d_table + (size_t)t * n * g.table;
    ++g_kernel_launches, k_bases_to_niels<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cur, n, g.table, tbl);
    ++g_kernel_launches, k_table_fill<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(n, g.table, tbl);
    ++g_kernel_launches, k_table_halve_first<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, g.table, tbl);
  }
}

// ------------------------------------------------------------------------------------------------ recode
__device__ __forceinline__ uint32_t recode_one(const fl_t *src, const MsmGeom &g, uint16_t *dst, size_t plane, uint32_t *used) {
  *used = 0;
  if (!src) {
    for (int w = 0; w < g.windows; w++) dst[(size_t)w * plane] = 0;
    return 0;
  }
  const uint4 *q = reinterpret_cast<const uint4 *>(src);
  uint4 lo = __ldg(q), hi = __ldg(q + 1);
  fl_t x;
  x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w; x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
  return msm_recode_value(x, g, dst, plane, used);
}
// Grid-stride: a block recodes several 256-scalar tiles and reports its window mask and its non-zero count ONCE. (One atomic
// per warp - 524 288 same-address atomics for CNN A's comb_ops commitment - kept this kernel at 1.1 TB/s: the L2 serialises them.)
__global__ void __launch_bounds__(256) k_recode(const fl_t *scalars, size_t rows, size_t cols, size_t ld, const fl_t *extra, MsmGeom g,
                                                size_t stride, uint16_t *digits, unsigned long long *nonzero, uint32_t *wmask) {
  uint32_t nz = 0, used = 0;
  const size_t total = rows * stride, step = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    size_t row = idx / stride, col = idx % stride;
    const fl_t *src = col < cols ? scalars + row * ld + col : (col == cols && extra ? extra + row : nullptr);
    uint32_t u = 0;
    nz += recode_one(src, g, digits + row * stride + col, total, &u);
    used |= u;
  }
  if (!wmask && !nonzero) return;
  __shared__ uint32_t s_nz[8], s_used[8];
  nz = __reduce_add_sync(0xffffffffu, nz);
  used = __reduce_or_sync(0xffffffffu, used);
  if ((threadIdx.x & 31) == 0) { s_nz[threadIdx.x >> 5] = nz; s_used[threadIdx.x >> 5] = used; }
  __syncthreads();
  if (threadIdx.x == 0) {
    nz = used = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) { nz += s_nz[k]; used |= s_used[k]; }
    if (wmask && used) atomicOr(wmask, used);  // windows that hold a non-zero digit anywhere in this launch (the accumulate kernel skips the others)
    if (nonzero && nz) atomicAdd(nonzero, (unsigned long long)nz);
  }
}
void launch_recode(const fl_t *d_scalars, size_t rows, size_t cols, size_t ld, const fl_t *d_extra, const MsmGeom &g, uint16_t *d_digits,
                   unsigned long long *d_nonzero, cudaStream_t st, uint32_t *d_wmask) {
  size_t stride = msm_col_stride(cols + (d_extra ? 1 : 0));
  size_t total = rows * stride;
  if (d_wmask) cudaMemsetAsync(d_wmask, 0, sizeof(uint32_t), st);
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)148 * 32;  // 4 resident blocks per SM x 8 waves
  if (blocks > cap) blocks = cap;
  ++g_kernel_launches, k_recode<<<(unsigned)blocks, 256, 0, st>>>(d_scalars, rows, cols, ld, d_extra, g, stride, d_digits, d_nonzero, d_wmask);
}

// ------------------------------------------------------------------------------------------------ accumulate
// Two interchangeable accumulators for the hot loop (same table, same group element, hence the same bytes):
//   Acc9  : extended coordinates in nine 29-bit limbs each, carry-free multiplier of fp29.cuh (kPolicy: where the 64-bit
//           accumulations of the partial products run, see f9_mul)
//   Acc8  : ed.cuh's 8 x u32 limbs with carry-chained IMAD.WIDE rows (the round-1 kernel)
// Table entries are HALVED ((y+x)/2, (y-x)/2, d x y): Hisil-Wong-Carter-Dawson mixed addition with Z2 = 1 and D = Z1 (instead
// of 2 Z1), seven multiplications; the sign of the digit swaps the two multiplicands (by address) and negates T.
template <int kPolicy>
struct Acc9 {
  f9 X, Y, Z, T;
  __device__ __forceinline__ void init() { X = f9_zero(); Y = f9_one(); Z = f9_one(); T = f9_zero(); }
  __device__ __forceinline__ ge_t get() const { ge_t r; r.X = f9_to_fp(X); r.Y = f9_to_fp(Y); r.Z = f9_to_fp(Z); r.T = f9_to_fp(T); return r; }
  // operand bounds as stated in fp29.cuh: E, F are differences of two normal values (signed limbs), G, H sums (unsigned product)
  __device__ __forceinline__ void madd(const niels_t *e, uint32_t neg) {
    const fp_t *p_yp = neg ? &e->ym : &e->yp, *p_ym = neg ? &e->yp : &e->ym;
    fp_t wyp = ldg_fp(p_yp), wym = ldg_fp(p_ym), wt = ldg_fp(&e->t2d);
    f9 a = f9_mul<false, kPolicy>(f9_sub(Y, X), f9_from_fp(wym));
    f9 b = f9_mul<false, kPolicy>(f9_add(Y, X), f9_from_fp(wyp));
    f9 c = f9_mul<false, kPolicy>(f9_cneg(T, neg != 0), f9_from_fp(wt));
    f9 e_ = f9_sub(b, a), h = f9_add(b, a), f = f9_sub(Z, c), g = f9_add(Z, c);
    X = f9_mul<false, kPolicy>(e_, f); Y = f9_mul<true, kPolicy>(g, h); Z = f9_mul<false, kPolicy>(f, g); T = f9_mul<false, kPolicy>(e_, h);
  }
};
template <uint32_t kAluRows>
struct Acc8 {
  ge_t p;
  __device__ __forceinline__ void init() { p = ge_identity(); }
  __device__ __forceinline__ ge_t get() const { return p; }
  __device__ __forceinline__ void madd(const niels_t *e, uint32_t neg) {
    const fp_t *p_yp = neg ? &e->ym : &e->yp, *p_ym = neg ? &e->yp : &e->ym;
    fp_t yp = ldg_fp(p_yp), ym = ldg_fp(p_ym), t2d = ldg_fp(&e->t2d);
    fp_t a = fp_mul_p<kAluRows>(fp_sub(p.Y, p.X), ym);
    fp_t b = fp_mul_p<kAluRows>(fp_add(p.Y, p.X), yp);
    fp_t c = fp_mul_p<kAluRows>(p.T, t2d);
    fp_t s1 = fp_sub(p.Z, c), s2 = fp_add(p.Z, c);
    fp_t f, g;
#pragma unroll
    for (int i = 0; i < 8; i++) { f.v[i] = neg ? s2.v[i] : s1.v[i]; g.v[i] = neg ? s1.v[i] : s2.v[i]; }
    fp_t e_ = fp_sub(b, a), h = fp_add(b, a);
    p.X = fp_mul_p<kAluRows>(e_, f); p.Y = fp_mul_p<kAluRows>(g, h); p.Z = fp_mul_p<kAluRows>(f, g); p.T = fp_mul_p<kAluRows>(e_, h);
  }
};
// grid (ceil(rows / 128), kMsmGroup, segs), block 128: thread = (row, local window w', column segment)
// kPrefetch: 1 = the table entries of the NEXT column are requested into L2 (prefetch.global.L2) while the current column's
// additions run: the entry address depends on a digit load, and that two-step chain to DRAM is what the loop's long-scoreboard
// stalls wait for (2.7 of 13.4 stall cycles per issue in profiles/r2_ncu_msm_accumulate_summary.csv)
template <class Acc, int kPrefetch = 0>
__device__ __forceinline__ void msm_accumulate_body(const niels_t *table, const MsmGeom &g, const uint16_t *digits, size_t rows, size_t cols,
                                                    size_t cols_total, size_t extra_base, size_t stride, size_t n_bases, size_t seg_len,
                                                    ge_t *partial, const uint32_t *wmask) {
  size_t row = (size_t)blockIdx.x * kMsmRowsPerBlock + threadIdx.x;
  if (row >= rows) return;
  const int wl = blockIdx.y;
  const size_t seg = blockIdx.z, segs = gridDim.z;
  // a local window none of whose sub-table windows holds a non-zero digit anywhere (small scalars: addresses, timestamps,
  // 128-bit weights) is finished at once; its digits are not even read
  {
    uint32_t live = wmask ? *wmask : 0xffffffffu, mine = 0;
    for (int t = 0; t < g.sub; t++) mine |= (t * g.group + wl) < g.windows ? (live >> (t * g.group + wl)) & 1u : 0u;
    if (!mine) {
      st_ge(partial + (row * g.group + wl) * segs + seg, ge_identity());
      return;
    }
  }
  size_t c0 = seg * seg_len, c1 = c0 + seg_len < cols_total ? c0 + seg_len : cols_total;
  const size_t plane = rows * stride;
  const uint16_t *dg = digits + row * stride;
  Acc acc;
  acc.init();
  for (size_t col = c0; col < c1; col++) {
    size_t base = col < cols ? col : extra_base;
    if (kPrefetch && col + 1 < c1) {
      size_t nbase = col + 1 < cols ? col + 1 : extra_base;
#pragma unroll
      for (int t = 0; t < g.sub; t++) {
        int w = t * g.group + wl;
        if (w < g.windows) {
          uint32_t d = dg[(size_t)w * plane + col + 1];
          if (d) {
            const char *e = reinterpret_cast<const char *>(table + ((size_t)t * n_bases + nbase) * g.table + ((d & 0x7fffu) - 1u));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(e));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(e + 64));  // (a 96-byte entry at 32-byte alignment can straddle a 128-byte line)
          }
        }
      }
    }
#pragma unroll 1
    for (int t = 0; t < g.sub; t++) {  // not unrolled: one copy of the 7-multiplication body keeps the loop inside the I-cache
      int w = t * g.group + wl;
      if (w >= g.windows) break;
      uint32_t d = dg[(size_t)w * plane + col];
      if (d) acc.madd(table + ((size_t)t * n_bases + base) * g.table + ((d & 0x7fffu) - 1u), d >> 15);
    }
  }
  st_ge(partial + (row * g.group + wl) * segs + seg, acc.get());
}
#define VPIN_MSM_ACC_ARGS const niels_t *table, MsmGeom g, const uint16_t *digits, size_t rows, size_t cols, size_t cols_total, size_t extra_base, \
                          size_t stride, size_t n_bases, size_t seg_len, ge_t *partial, const uint32_t *wmask
#define VPIN_MSM_ACC_PASS table, g, digits, rows, cols, cols_total, extra_base, stride, n_bases, seg_len, partial, wmask
__global__ void __launch_bounds__(kMsmRowsPerBlock, 6) k_msm_accumulate(VPIN_MSM_ACC_ARGS) { msm_accumulate_body<Acc8<0>>(VPIN_MSM_ACC_PASS); }
__global__ void __launch_bounds__(kMsmRowsPerBlock, 6) k_msm_accumulate_pf(VPIN_MSM_ACC_ARGS) { msm_accumulate_body<Acc8<0>, 1>(VPIN_MSM_ACC_PASS); }
// 0x8888: the odd-column products of rows 1, 3, 5, 7 accumulate on the ALU pipe (the best of the row patterns tried)
__global__ void __launch_bounds__(kMsmRowsPerBlock, 5) k_msm_accumulate_a2(VPIN_MSM_ACC_ARGS) { msm_accumulate_body<Acc8<0x8888u>>(VPIN_MSM_ACC_PASS); }
__global__ void __launch_bounds__(kMsmRowsPerBlock, 4) k_msm_accumulate_f9p0(VPIN_MSM_ACC_ARGS) { msm_accumulate_body<Acc9<0>>(VPIN_MSM_ACC_PASS); }
__global__ void __launch_bounds__(kMsmRowsPerBlock, 4) k_msm_accumulate_f9p1(VPIN_MSM_ACC_ARGS) { msm_accumulate_body<Acc9<1>>(VPIN_MSM_ACC_PASS); }
// VPIN_MSM_VARIANT selects one of the measured alternatives of the hot loop (all bit-identical; 2^22 uniform scalars on a B200,
// profiles/r2_msm_variants.log): 0 (default) 8 x 32 carry-chained 5.01 ms | 1 radix 2^29, accumulation in the multiplier 6.38 ms |
// 2 radix 2^29, accumulation on the ALU pipe 7.31 ms | 12 8 x 32 with four rows accumulated on the ALU pipe 5.21 ms.
// (One level of subtractive Karatsuba on the 8 x 8 limb product - 48 + 8 instead of 64 + 8 multiplies - was also built: ptxas
// places the ~60 extra additions and moves on the multiply pipe as IMAD.X / IMAD.MOV, 5.5 - 5.8 ms; dropped.)
static int msm_variant() {
  static const int v = [] { const char *e = getenv("VPIN_MSM_VARIANT"); return e ? atoi(e) : 0; }();
  return v;
}
__device__ __forceinline__ ge_t shfl_down_ge(const ge_t &g, int off) {
  ge_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_down_sync(0xffffffffu, g.X.v[i], off);
    r.Y.v[i] = __shfl_down_sync(0xffffffffu, g.Y.v[i], off);
    r.Z.v[i] = __shfl_down_sync(0xffffffffu, g.Z.v[i], off);
    r.T.v[i] = __shfl_down_sync(0xffffffffu, g.T.v[i], off);
  }
  return r;
}
__device__ __forceinline__ fp_t quad_from(const fp_t &v, int src) {
  fp_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src, 4);
  return r;
}
__device__ __forceinline__ fp_t quad_xor1(const fp_t &v) {
  fp_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, v.v[i], 1, 4);
  return r;
}
__device__ __forceinline__ fp_t fp_sel(bool c, const fp_t &a, const fp_t &b) {
  fp_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}
// second round of a doubling / addition: lane 0 E F (X3), lane 1 G H (Y3), lane 2 F G (Z3), lane 3 E H (T3)
__device__ __forceinline__ fp_t quad_efgh(int q, const fp_t &e, const fp_t &f, const fp_t &g, const fp_t &h) {
  fp_t m1 = fp_sel(q == 1, g, fp_sel(q == 2, f, e));
  fp_t m2 = fp_sel(q == 0, f, fp_sel(q == 2, g, h));
  return fp_mul(m1, m2);
}
// c: coordinate q of the point (lane 0 X, 1 Y, 2 Z, 3 T) -> coordinate q of its double
__device__ __forceinline__ fp_t quad_dbl(int q, const fp_t &c) {
  fp_t x = quad_from(c, 0), y = quad_from(c, 1);
  fp_t s = fp_sqr(fp_sel(q == 3, fp_add(x, y), c));
  fp_t a = quad_from(s, 0), b = quad_from(s, 1), cc = quad_from(s, 2), xy2 = quad_from(s, 3);
  cc = fp_add(cc, cc);
  fp_t e = fp_sub(fp_sub(xy2, a), b), g = fp_sub(b, a), f = fp_sub(g, cc), h = fp_sub(fp_neg(a), b);
  return quad_efgh(q, e, f, g, h);
}
// coordinate q of P + Q; Q is read from memory (all lanes see the same point), t2d = Q.T * 2d
__device__ __forceinline__ fp_t quad_add(int q, const fp_t &c, const ge_t &Q, const fp_t &t2d) {
  fp_t o = quad_xor1(c);  // lane 0: Y1, lane 1: X1, lane 2: T1, lane 3: Z1
  fp_t m1 = fp_sel(q == 0, fp_sub(o, c), fp_sel(q == 1, fp_add(c, o), o));
  fp_t m2 = fp_sel(q == 0, fp_sub(Q.Y, Q.X), fp_sel(q == 1, fp_add(Q.Y, Q.X), fp_sel(q == 2, t2d, Q.Z)));
  fp_t s = fp_mul(m1, m2);
  fp_t a = quad_from(s, 0), b = quad_from(s, 1), cc = quad_from(s, 2), d = quad_from(s, 3);
  d = fp_add(d, d);
  fp_t e = fp_sub(b, a), f = fp_sub(d, cc), g = fp_add(d, cc), h = fp_add(b, a);
  return quad_efgh(q, e, f, g, h);
}
__device__ __forceinline__ ge_t shfl_ge(const ge_t &g, int src) {
  ge_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_sync(0xffffffffu, g.X.v[i], src);
    r.Y.v[i] = __shfl_sync(0xffffffffu, g.Y.v[i], src);
    r.Z.v[i] = __shfl_sync(0xffffffffu, g.Z.v[i], src);
    r.T.v[i] = __shfl_sync(0xffffffffu, g.T.v[i], src);
  }
  return r;
}
// Sum of the points held by each group of `lanes` consecutive lanes (a power of two), left in the group's first lane - the
// shuffle tree of the small-row MSM and of the segment sum, with every addition done by FOUR lanes (the quad of lanes
// 4k .. 4k + 3 takes the k-th addition of a pass): three multiplications deep instead of nine, and a level of fewer than eight
// additions no longer costs a warp-wide one. A bullet-reduction round is 9 - 11 such levels on the proof's critical path.
// Every lane of the warp must call it; lanes without data hold the identity.
__device__ __forceinline__ ge_t warp_tree_sum_quad(ge_t acc, int lanes) {
  const int lane = threadIdx.x & 31, q = lane & 3, quad = lane >> 2;
  for (int s = lanes >> 1; s > 0; s >>= 1) {
    const int adds = (32 / lanes) * s;  // additions of this level: (a, a + s) with a = group * lanes + i, i < s
    for (int p0 = 0; p0 < adds; p0 += 8) {
      int j = p0 + quad;
      if (j >= adds) j = adds - 1;  // (idle quads repeat the last addition; nobody takes their result)
      const int a = (j / s) * lanes + (j % s);
      ge_t P = shfl_ge(acc, a), Q = shfl_ge(acc, a + s);
      fp_t m1 = fp_sel(q == 0, fp_sub(P.Y, P.X), fp_sel(q == 1, fp_add(P.Y, P.X), fp_sel(q == 2, P.T, P.Z)));
      fp_t m2 = fp_sel(q == 0, fp_sub(Q.Y, Q.X), fp_sel(q == 1, fp_add(Q.Y, Q.X), fp_sel(q == 2, Q.T, Q.Z)));
      fp_t sres = fp_mul(m1, m2);
      fp_t sc = fp_mul(sres, fp_d2());  // lane 2: C = 2 d T1 T2
      sres = fp_sel(q == 2, sc, sres);
      fp_t a_ = quad_from(sres, 0), b_ = quad_from(sres, 1), c_ = quad_from(sres, 2), d_ = quad_from(sres, 3);
      d_ = fp_add(d_, d_);
      fp_t e = fp_sub(b_, a_), f = fp_sub(d_, c_), gg = fp_add(d_, c_), h = fp_add(b_, a_);
      fp_t r = quad_efgh(q, e, f, gg, h);  // lane q: coordinate q of P + Q
      // the lane that holds the first operand takes the whole sum: addition j of this pass sits in quad j - p0
      const int i = lane % lanes, mine = (lane / lanes) * s + i;
      const bool dest = i < s && mine >= p0 && mine < p0 + 8 && mine < adds;
      const int src = dest ? 4 * (mine - p0) : lane & ~3;
      ge_t sum;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        sum.X.v[k] = __shfl_sync(0xffffffffu, r.v[k], src);
        sum.Y.v[k] = __shfl_sync(0xffffffffu, r.v[k], src + 1);
        sum.Z.v[k] = __shfl_sync(0xffffffffu, r.v[k], src + 2);
        sum.T.v[k] = __shfl_sync(0xffffffffu, r.v[k], src + 3);
      }
      if (dest) acc = sum;
    }
  }
  return acc;
}
// Few rows (the bullet-reduction L / R rows, single Pedersen commitments): rows cannot fill a warp, so the lanes of a warp
// take the COLUMNS of one (row, local window, segment) instead and the 32 partial sums are added by a shuffle tree.
// grid (segs, kMsmGroup, rows), one warp per block.
__global__ void __launch_bounds__(32) k_msm_accumulate_small(const niels_t *table, MsmGeom g, const uint16_t *digits, size_t rows, size_t cols,
                                                             size_t cols_total, size_t extra_base, size_t stride, size_t n_bases,
                                                             size_t seg_len, ge_t *partial, int quad_tree) {
  const size_t row = blockIdx.z, seg = blockIdx.x, segs = gridDim.x;
  const int wl = blockIdx.y, lane = threadIdx.x;
  size_t c0 = seg * seg_len, c1 = c0 + seg_len < cols_total ? c0 + seg_len : cols_total;
  const size_t plane = rows * stride;
  const uint16_t *dg = digits + row * stride;
  Acc8<0> a8;
  a8.init();
  for (size_t col = c0 + lane; col < c1; col += 32) {
    size_t base = col < cols ? col : extra_base;
#pragma unroll 1
    for (int t = 0; t < g.sub; t++) {
      int w = t * g.group + wl;
      if (w >= g.windows) break;
      uint32_t d = dg[(size_t)w * plane + col];
      if (d) a8.madd(table + ((size_t)t * n_bases + base) * g.table + ((d & 0x7fffu) - 1u), d >> 15);
    }
  }
  ge_t acc = quad_tree ? warp_tree_sum_quad(a8.get(), 32) : a8.get();
  if (!quad_tree) {
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) {
      ge_t o = shfl_down_ge(acc, off);
      acc = ge_add(acc, o);
    }
  }
  if (lane == 0) st_ge(partial + (row * g.group + wl) * segs + seg, acc);
}
static bool tree_quad() {  // VPIN_TREE_QUAD=0: one lane per addition in the shuffle trees (for comparison)
  static const bool v = [] { const char *e = getenv("VPIN_TREE_QUAD"); return !e || atoi(e) != 0; }();
  return v;
}
static const size_t kMsmSmallRows = 16;
size_t msm_num_segments(size_t rows, size_t cols_total, const MsmGeom &g) {
  if (rows <= kMsmSmallRows) {
    const size_t want_warps = (size_t)148 * 16;
    size_t per = rows * g.group;
    size_t segs = (want_warps + per - 1) / per;
    size_t max_segs = (cols_total + 63) / 64;  // at least two columns per lane
    if (segs > max_segs) segs = max_segs;
    return segs < 1 ? 1 : segs;
  }
  const size_t want_threads = (size_t)148 * 6144;  // ~6 waves of blocks: short blocks balance the SMs (measured: 610 -> 736 Mpoints/s from 1024)
  size_t per = rows * g.group;
  size_t segs = (want_threads + per - 1) / per;
  size_t max_segs = (cols_total + 7) / 8;  // at least 8 columns per thread
  if (segs > max_segs) segs = max_segs;
  if (segs < 1) segs = 1;
  if (segs > 65535) segs = 65535;
  return segs;
}
void launch_msm_accumulate(const MsmTable &t, const uint16_t *d_digits, size_t rows, size_t cols, bool has_extra, size_t extra_base,
                           size_t segs, ge_t *d_partial, cudaStream_t st, const uint32_t *d_wmask, int blocks_per_sm) {
  size_t cols_total = cols + (has_extra ? 1 : 0);
  size_t stride = msm_col_stride(cols_total);
  size_t seg_len = (cols_total + segs - 1) / segs;
  if (rows <= kMsmSmallRows) {
    dim3 grid((unsigned)segs, t.geom.group, (unsigned)rows);
    ++g_kernel_launches, k_msm_accumulate_small<<<grid, 32, 0, st>>>(t.d_table, t.geom, d_digits, rows, cols, cols_total, extra_base, stride,
                                                                      t.n_bases, seg_len, d_partial, tree_quad() ? 1 : 0);
    return;
  }
  dim3 grid((unsigned)((rows + kMsmRowsPerBlock - 1) / kMsmRowsPerBlock), t.geom.group, (unsigned)segs);
  ++g_kernel_launches;
  if (blocks_per_sm > 0 && blocks_per_sm < 6) {  // occupancy cap: 227 KB of shared memory per SM / blocks_per_sm, minus a margin
    static const bool ok = cudaFuncSetAttribute(k_msm_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
    size_t smem = ok ? (size_t)(220 * 1024) / (blocks_per_sm + 1) + 1024 : 0;  // more than a (cap + 1)-th of the SM: cap + 1 blocks cannot fit
    if (smem > 200 * 1024) smem = 200 * 1024;
    k_msm_accumulate<<<grid, kMsmRowsPerBlock, smem, st>>>(t.d_table, t.geom, d_digits, rows, cols, cols_total, extra_base, stride, t.n_bases, seg_len,
                                                          d_partial, d_wmask);
    return;
  }
#define VPIN_MSM_LAUNCH(K) K<<<grid, kMsmRowsPerBlock, 0, st>>>(t.d_table, t.geom, d_digits, rows, cols, cols_total, extra_base, stride, t.n_bases, seg_len, d_partial, d_wmask)
  switch (msm_variant()) {
    case 1: VPIN_MSM_LAUNCH(k_msm_accumulate_f9p0); break;
    case 2: VPIN_MSM_LAUNCH(k_msm_accumulate_f9p1); break;
    case 12: VPIN_MSM_LAUNCH(k_msm_accumulate_a2); break;
    case 20: VPIN_MSM_LAUNCH(k_msm_accumulate_pf); break;
    default: VPIN_MSM_LAUNCH(k_msm_accumulate); break;
  }
#undef VPIN_MSM_LAUNCH
}

// finish, step 1 (only when a row was split into segments): `lanes` lanes per (row, local window) add the segment partials
// (lane-strided, then a shuffle tree) -> sums[row][w']. The kernel is bound by the multiply pipe, and a warp-wide addition costs
// the same whether 1 or 32 of its lanes hold data: with a whole warp per pair the tree's five levels were 5 of the 6 warp-wide
// additions a pair of 45 segments cost (23 % of the lanes useful, 344 us for CNN A's comb_ops commitment). `lanes` now comes from
// a small cost model (msm_segsum_lanes): 2 - 4 lanes for the 20480 pairs x 45 segments of that commitment, a full tree for the
// 10 pairs of a bullet-reduction round.
// publish: optional {counter, host-mapped sequence word, value}: the last block to finish stores the value there (system scope),
// which tells the host that every window sum has landed in its mapped slot — saves the separate one-thread launch
struct SegsumPublish { unsigned *counter; volatile uint32_t *seq_word; uint32_t seq; };
__global__ void __launch_bounds__(128) k_msm_segsum(const ge_t *partial, size_t pairs, size_t segs, int lanes, int quad_tree, ge_t *sums,
                                                    SegsumPublish pub) {
  const int lane = threadIdx.x & 31, sub = lane & (lanes - 1);
  const size_t warp = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const size_t pair = warp * (32 / lanes) + lane / lanes;  // (row, w') flattened
  const bool live = pair < pairs;
  {
    const ge_t *p = partial + (live ? pair : 0) * segs;
    ge_t acc = ge_identity();
    bool first = true;
    if (live)
      for (size_t s = sub; s < segs; s += lanes) {
        ge_t q = ld_ge(p + s);
        acc = first ? q : ge_add(acc, q);
        first = false;
      }
    if (quad_tree) acc = warp_tree_sum_quad(acc, lanes);
    else
      for (int off = lanes >> 1; off > 0; off >>= 1) {  // (lanes beyond the data hold the identity; a read across the group's end
        ge_t o = shfl_down_ge(acc, off);               //  only reaches lanes whose sums nobody uses)
        acc = ge_add(acc, o);
      }
    if (live && sub == 0) st_ge(sums + pair, acc);
  }
  if (pub.counter) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(pub.counter, 1u) == gridDim.x - 1) {
      *pub.counter = 0;
      __threadfence_system();
      *pub.seq_word = pub.seq;
    }
  }
}
// lanes per pair: the power of two that minimises (multiplications in sequence) x max(latency of one, multiply-pipe time of one
// for all warps) with ceil(segs / lanes) - 1 serial additions of nine multiplications and the passes of the quad tree
// (warp_tree_sum_quad: three in sequence, four executed). Few pairs (the two rows of a bullet-reduction round) are latency-bound
// and get a wide tree; thousands of pairs are pipe-bound and get few lanes.
static int msm_segsum_lanes(size_t pairs, size_t segs) {
  static const int forced = [] { const char *e = getenv("VPIN_SEGSUM_LANES"); return e ? atoi(e) : 0; }();  // (experiments: 1, 2, .., 32)
  if (forced >= 1 && forced <= 32 && (forced & (forced - 1)) == 0) return forced;
  const double mul_latency = 650.0;          // cycles of one dependent F_p multiplication in a lone warp
  const double pipe_per_warp_mul = 288.0 / 592;  // 72 IMAD.WIDE x 4 cycles, spread over 148 x 4 sub-partitions
  const bool quad = tree_quad();
  int best = 1;
  double best_cost = 0;
  for (int lanes = 1; lanes <= 32; lanes <<= 1) {
    double serial = (double)((segs + lanes - 1) / lanes) - 1;
    double seq = serial * 9, work = serial * 9;
    for (int s = lanes >> 1; s > 0; s >>= 1) {
      if (quad) {
        int passes = ((32 / lanes) * s + 7) / 8;
        seq += 3.0 * passes;
        work += 4.0 * passes;
      } else {
        seq += 9;
        work += 9;
      }
    }
    if (seq < 1) seq = 1;
    double warps = (double)pairs * lanes / 32.0;
    double per_mul = warps * pipe_per_warp_mul > mul_latency ? warps * pipe_per_warp_mul : mul_latency;
    double cost = (warps * pipe_per_warp_mul > mul_latency ? work : seq) * per_mul;  // pipe-bound: what is executed counts
    if (lanes == 1 || cost < best_cost) { best = lanes; best_cost = cost; }
    if ((size_t)lanes >= segs) break;
  }
  return best;
}
static unsigned msm_segsum_blocks(size_t pairs, int lanes) {
  size_t per_block = (size_t)4 * (32 / lanes);
  return (unsigned)((pairs + per_block - 1) / per_block);
}
// finish, step 2: one thread per row runs the Horner pass over the kMsmGroup window sums and encodes the point
__global__ void __launch_bounds__(32) k_msm_horner(const ge_t *sums, size_t rows, MsmGeom g, ge_t *out, uint8_t *comp) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const ge_t *p = sums + row * g.group;
  ge_t h = ld_ge(p + g.group - 1);
  for (int w = g.group - 2; w >= 0; w--) {
    for (int i = 0; i < g.W; i++) h = ge_dbl(h);
    h = ge_add(h, ld_ge(p + w));
  }
  if (out) st_ge(out + row, h);
  if (comp) {
    uint8_t b[32];
    ge_compress(h, b);
    uint4 *q = reinterpret_cast<uint4 *>(comp + 32 * row);
    uint32_t w[8];
    for (int k = 0; k < 8; k++) w[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24);
    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
    q[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
// The same pass with FOUR lanes per row. A commitment ends with this kernel and the host waits for its bytes, so what counts
// is the length of the dependent chain, not the work: a doubling is two rounds of four independent multiplications (the
// squares of X, Y, Z, X + Y, then E F, G H, F G, E H), an addition likewise, and lane q of a quad computes the q-th product of
// each round; the products travel between the lanes by shuffles. 2 multiplications deep per doubling instead of 8, 3 per
// addition instead of 9. The encoding's inverse square root is a chain of squarings that cannot be split; lane 0 runs it.
__global__ void __launch_bounds__(128) k_msm_horner_quad(const ge_t *sums, size_t rows, MsmGeom g, ge_t *out, uint8_t *comp) {
  const int q = threadIdx.x & 3;
  size_t row = (size_t)blockIdx.x * 32 + (threadIdx.x >> 2);
  const bool live = row < rows;
  if (!live) row = rows - 1;  // (whole quads stay in step for the shuffles)
  const ge_t *p = sums + row * g.group;
  const fp_t *first = &p[g.group - 1].X;
  fp_t c = ld_fp(first + q);
  for (int w = g.group - 2; w >= 0; w--) {
    ge_t Q = ld_ge(p + w);
    fp_t t2d = fp_mul(Q.T, fp_d2());
#pragma unroll 1
    for (int i = 0; i < g.W; i++) c = quad_dbl(q, c);
    c = quad_add(q, c, Q, t2d);
  }
  ge_t h;
  h.X = quad_from(c, 0); h.Y = quad_from(c, 1); h.Z = quad_from(c, 2); h.T = quad_from(c, 3);
  if (!live || q != 0) return;
  if (out) st_ge(out + row, h);
  if (comp) {
    uint8_t b[32];
    ge_compress(h, b);
    uint4 *o4 = reinterpret_cast<uint4 *>(comp + 32 * row);
    uint32_t w[8];
    for (int k = 0; k < 8; k++) w[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24);
    o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o4[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
static bool horner_quad() {  // VPIN_HORNER_QUAD=0: the one-thread-per-row pass (for comparison)
  static const bool v = [] { const char *e = getenv("VPIN_HORNER_QUAD"); return !e || atoi(e) != 0; }();
  return v;
}
static void launch_horner(const ge_t *sums, size_t rows, const MsmGeom &g, ge_t *d_out, uint8_t *d_comp, cudaStream_t st) {
  ++g_kernel_launches;
  if (horner_quad()) k_msm_horner_quad<<<(unsigned)((rows + 31) / 32), 128, 0, st>>>(sums, rows, g, d_out, d_comp);
  else k_msm_horner<<<(unsigned)((rows + 31) / 32), 32, 0, st>>>(sums, rows, g, d_out, d_comp);
}
void launch_msm_segsum(const ge_t *d_partial, size_t rows, size_t segs, const MsmGeom &g, ge_t *d_sums, cudaStream_t st, unsigned *d_counter,
                       uint32_t *d_seq_word, uint32_t seq) {
  size_t pairs = rows * g.group;
  SegsumPublish pub{d_counter, d_seq_word, seq};
  const int lanes = msm_segsum_lanes(pairs, segs);
  ++g_kernel_launches, k_msm_segsum<<<msm_segsum_blocks(pairs, lanes), 128, 0, st>>>(d_partial, pairs, segs, lanes, tree_quad() ? 1 : 0, d_sums, pub);
}
void launch_msm_finish(const ge_t *d_partial, size_t rows, size_t segs, const MsmGeom &g, ge_t *d_sums, ge_t *d_out, uint8_t *d_comp,
                       cudaStream_t st) {
  const ge_t *sums = d_partial;
  if (segs > 1) {
    size_t pairs = rows * g.group;
    const int lanes = msm_segsum_lanes(pairs, segs);
    ++g_kernel_launches, k_msm_segsum<<<msm_segsum_blocks(pairs, lanes), 128, 0, st>>>(d_partial, pairs, segs, lanes, tree_quad() ? 1 : 0, d_sums, SegsumPublish{nullptr, nullptr, 0});
    sums = d_sums;
  }
  launch_horner(sums, rows, g, d_out, d_comp, st);
}

// ------------------------------------------------------------------------------------------------ encodings
__global__ void __launch_bounds__(64) k_compress(const ge_t *pts, size_t n, uint8_t *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[32];
  ge_compress(ld_ge(pts + i), b);
  uint4 *q = reinterpret_cast<uint4 *>(out + 32 * i);
  uint32_t w[8];
  for (int k = 0; k < 8; k++) w[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24);
  q[0] = make_uint4(w[0], w[1], w[2], w[3]);
  q[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
void launch_compress(const ge_t *d_pts, size_t n, uint8_t *d_out, cudaStream_t st) {
  ++g_kernel_launches, k_compress<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(d_pts, n, d_out);
}
__global__ void __launch_bounds__(64) k_decompress(const uint8_t *in, size_t n, ge_t *pts, uint8_t *ok) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[32];
  for (int k = 0; k < 32; k++) b[k] = in[32 * i + k];
  ge_t g;
  bool good = ge_decompress(b, &g);
  if (!good) g = ge_identity();
  st_ge(pts + i, g);
  if (ok) ok[i] = good ? 1 : 0;
}
void launch_decompress(const uint8_t *d_in, size_t n, ge_t *d_pts, uint8_t *d_ok, cudaStream_t st) {
  ++g_kernel_launches, k_decompress<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(d_in, n, d_pts, d_ok);
}
__global__ void __launch_bounds__(128) k_points_add(const ge_t *a, const ge_t *b, size_t n, ge_t *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_ge(out + i, ge_add(ld_ge(a + i), ld_ge(b + i)));
}
void launch_points_add(const ge_t *a, const ge_t *b, size_t n, ge_t *out, cudaStream_t st) {
  ++g_kernel_launches, k_points_add<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a, b, n, out);
}
__global__ void __launch_bounds__(64) k_from_uniform(const uint8_t *in, size_t n, ge_t *pts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[64];
  for (int k = 0; k < 64; k++) b[k] = in[64 * i + k];
  st_ge(pts + i, ge_from_uniform_bytes(b));
}
void launch_from_uniform_bytes(const uint8_t *d_in, size_t n, ge_t *d_pts, cudaStream_t st) {
  ++g_kernel_launches, k_from_uniform<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(d_in, n, d_pts);
}

}  // namespace vpin
