// Fused sumcheck round kernels for sm_100a: one launch per round binds every table with the previous challenge
// (Spartan/src/dense_mlpoly.rs:229-236) and evaluates the next round polynomial (Spartan/src/sumcheck.rs:287-357, :456-469,
// :619-652) in the same pass — each table element is read once and the half-length table written once per round
// (48 * k * L bytes for k tables of length L, the algorithmic figure of SURVEY.md 8d). The challenge travels as a kernel
// parameter; the round sums are reduced by the last block to arrive and stored straight into host-mapped pinned memory
// followed by a sequence number, so a round costs one launch and no memcpy / stream synchronisation.
//
// Layout of a bind+evaluate step on a table of current length 4q: thread i < q owns T[i], T[i+q], T[i+2q], T[i+3q];
//   lo' = T[i]   + r (T[i+2q] - T[i])      (element i       of the bound table, length 2q)
//   hi' = T[i+q] + r (T[i+3q] - T[i+q])    (element i + q   of the bound table)
// are stored in place (no other thread touches these four slots) and (lo', hi') is the pair the evaluation needs.
#include "launch_count.hpp"
#include <atomic>
#include <cstdlib>

#include "kernels_poly.cuh"
#include "msm_recode.cuh"

namespace vpin {


namespace {

__device__ __forceinline__ fl_t ldr(const fl_t *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1];
  fl_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ fl_t ldcg(const fl_t *p) {  // L2 (partials written by other blocks)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldcg(q), b = __ldcg(q + 1);
  fl_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void str(fl_t *p, const fl_t &x) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ fl_t shfl_down(const fl_t &x, int off) {
  fl_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], off);
  return r;
}
// block-wide sums of K accumulators; valid in thread 0
template <int K>
__device__ __forceinline__ void block_sum(fl_t (&acc)[K]) {
  __shared__ fl_t sm[K][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = fl_add(acc[k], shfl_down(acc[k], off));
  __syncthreads();  // sm may still be read by a previous call
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; k++) sm[k][warp] = acc[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      fl_t v = lane < nwarps ? sm[k][lane] : fl_zero();
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v = fl_add(v, shfl_down(v, off));
      acc[k] = v;
    }
  }
}
// Writes the block's K sums to partials[(inst * gridDim.x + blockIdx.x) * K ..]; the last block of instance `inst` to
// arrive adds all partials of the instance and stores them to slot->vals[inst * K ..]; the last instance to finish
// publishes the sequence number. counters[0] counts finished instances, counters[1 + inst] finished blocks of an instance.
template <int K>
__device__ __forceinline__ void publish(fl_t (&acc)[K], int inst, int ninst, fl_t *partials, unsigned *counters, RoundSlot *slot,
                                        uint32_t seq) {
  __shared__ bool is_last;
  block_sum<K>(acc);
  if (gridDim.x > 1) {  // (a single block per instance already holds the instance's sums)
    fl_t *mine = partials + ((size_t)inst * gridDim.x + blockIdx.x) * K;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < K; k++) str(mine + k, acc[k]);
      __threadfence();
      is_last = atomicAdd(counters + 1 + inst, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = fl_zero();
    const fl_t *p = partials + (size_t)inst * gridDim.x * K;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
#pragma unroll
      for (int k = 0; k < K; k++) acc[k] = fl_add(acc[k], ldcg(p + (size_t)b * K + k));
    block_sum<K>(acc);
    if (threadIdx.x == 0) counters[1 + inst] = 0;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) str(slot->vals + inst * K + k, acc[k]);
    __threadfence_system();
    if (atomicAdd(counters, 1u) == (unsigned)ninst - 1) {
      counters[0] = 0;
      __threadfence_system();
      *reinterpret_cast<volatile uint32_t *>(&slot->seq) = seq;
    }
  }
}

__device__ __forceinline__ fl_t bind1(const fl_t &lo, const fl_t &hi, const fl_t &r) { return fl_add(lo, fl_mul(r, fl_sub(hi, lo))); }

// ---- challenge mailbox (kernels_poly.cuh): the pre-launched kernel's wait for the challenge the host has not derived yet ----
// one 64-byte poll of a host-mapped slot; true when all eight (limb, tag) atoms carry `tag`
__device__ __forceinline__ bool chal_poll_host(const ChalSlot *slot, uint32_t tag, uint32_t (&v)[8]) {
  uint32_t t[8];
#pragma unroll
  for (int k = 0; k < 4; k++)
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[2 * k]), "=r"(t[2 * k]), "=r"(v[2 * k + 1]), "=r"(t[2 * k + 1])
                 : "l"(reinterpret_cast<const char *>(slot) + 16 * k)
                 : "memory");
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 8; k++) ok = ok && t[k] == tag;
  return ok;
}
// Replaces r by the posted challenge when the launch carries a mailbox reference. Block-uniform result; false = the mailbox
// was aborted and NO thread of the grid may publish anything (the host then sees a drained stream without a result).
__device__ __forceinline__ bool chal_fetch(const ChalRef &c, fl_t &r) {
  if (!c.slot) return true;
  __shared__ uint32_t s_r[8];
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int ok = 0;
    ChalLatch *L = c.latch;
    volatile uint32_t *owner = &L->owner, *ready = &L->ready, *abortw = &L->abort;
    const bool single = gridDim.x * gridDim.y * gridDim.z == 1;
    bool poller = single;
    if (!single) {  // the first block of the grid to arrive polls the host for everybody
      uint32_t old = *owner;
      while ((int32_t)(c.tag - old) > 0) {
        uint32_t prev = atomicCAS(&L->owner, old, c.tag);
        if (prev == old) { poller = true; break; }
        old = prev;
      }
    }
    const long long t0 = clock64(), limit = (long long)c.timeout_ms << 21;  // ~2.1 M cycles per millisecond
    if (poller) {
      for (;;) {
        if (*abortw) break;
        if (chal_poll_host(c.slot, c.tag, v)) { ok = 1; break; }
        if (clock64() - t0 > limit) { *abortw = 1; break; }
      }
      if (!single) {
        if (ok) {
          volatile uint32_t *dst = L->r;
#pragma unroll
          for (int k = 0; k < 8; k++) dst[k] = v[k];
        }
        __threadfence();
        *ready = c.tag;
      }
    } else {
      for (;;) {
        if (*ready == c.tag) {
          __threadfence();
          if (!*abortw) {
            const volatile uint32_t *src = L->r;
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = src[k];
            ok = 1;
          }
          break;
        }
        if (clock64() - t0 > 2 * limit) break;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) s_r[k] = v[k];
    s_ok = ok;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = s_r[k];
  return s_ok != 0;
}

// loads the evaluation pair (lo, hi) of thread item i; with kBind the table is first bound in place with r
template <bool kBind>
__device__ __forceinline__ void load_pair(fl_t *T, size_t i, size_t q, const fl_t &r, fl_t &lo, fl_t &hi) {
  if (kBind) {
    fl_t t0 = ldr(T + i), t1 = ldr(T + i + q), t2 = ldr(T + i + 2 * q), t3 = ldr(T + i + 3 * q);
    lo = bind1(t0, t2, r);
    hi = bind1(t1, t3, r);
    str(T + i, lo);
    str(T + i + q, hi);
  } else {
    lo = ldr(T + i);
    hi = ldr(T + i + q);
  }
}
// ---- A (B C - D) at t = 0, 2, 3 ----
template <bool kBind>
__global__ void __launch_bounds__(kRedThreads, 2) k_round_cubic_additive(fl_t *A, fl_t *B, fl_t *C, fl_t *D, size_t q, fl_t r, fl_t *partials,
                                                                      unsigned *counters, RoundSlot *slot, uint32_t seq) {
  fl_t acc[3] = {fl_zero(), fl_zero(), fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += stride) {
    fl_t a0, a1, b0, b1, c0, c1, d0, d1;
    load_pair<kBind>(A, i, q, r, a0, a1);
    load_pair<kBind>(B, i, q, r, b0, b1);
    load_pair<kBind>(C, i, q, r, c0, c1);
    load_pair<kBind>(D, i, q, r, d0, d1);
    acc[0] = fl_add(acc[0], fl_mul(a0, fl_sub(fl_mul(b0, c0), d0)));
    fl_t da = fl_sub(a1, a0), db = fl_sub(b1, b0), dc = fl_sub(c1, c0), dd = fl_sub(d1, d0);
    fl_t a2 = fl_add(a1, da), b2 = fl_add(b1, db), c2 = fl_add(c1, dc), d2 = fl_add(d1, dd);
    acc[1] = fl_add(acc[1], fl_mul(a2, fl_sub(fl_mul(b2, c2), d2)));
    fl_t a3 = fl_add(a2, da), b3 = fl_add(b2, db), c3 = fl_add(c2, dc), d3 = fl_add(d2, dd);
    acc[2] = fl_add(acc[2], fl_mul(a3, fl_sub(fl_mul(b3, c3), d3)));
  }
  publish<3>(acc, 0, 1, partials, counters, slot, seq);
}

// ---- the same round with A = eq(tau, .) factored out ----
// s_j(X) = sum_x eq(tau, (r_<j, X, x)) (B C - D)(r_<j, X, x) = E_j eq(tau_j, X) t_j(X) with E_j = prod_{k<j} eq(tau_k, r_k) and
// t_j(X) = sum_x eq(tau_{>j}, x) (B C - D)(r_<j, X, x), a QUADRATIC. The kernel binds B, C, D and returns t_j(0) and the leading
// coefficient t_j(inf) = sum_x eq_rest[x] (B1 - B0)(C1 - C0) (D is multilinear, it has no X^2 term); the host recovers t_j(1) from
// the running claim and forms the three evaluations the reference sends (prover.cu zk_sumcheck). The eq table is neither bound
// nor streamed four times per round: 10 multiplications and 13 element reads per thread item instead of 14 and 16.
template <bool kBind>
__global__ void __launch_bounds__(kRedThreads, 2) k_round_r1cs_split(const fl_t *E, fl_t *B, fl_t *C, fl_t *D, size_t q, fl_t r, fl_t *partials,
                                                                  unsigned *counters, RoundSlot *slot, uint32_t seq) {
  fl_t acc[2] = {fl_zero(), fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += stride) {
    fl_t b0, b1, c0, c1, d0, d1;
    load_pair<kBind>(B, i, q, r, b0, b1);
    load_pair<kBind>(C, i, q, r, c0, c1);
    load_pair<kBind>(D, i, q, r, d0, d1);
    fl_t e = ldr(E + i);
    acc[0] = fl_add(acc[0], fl_mul(e, fl_sub(fl_mul(b0, c0), d0)));
    acc[1] = fl_add(acc[1], fl_mul(e, fl_mul(fl_sub(b1, b0), fl_sub(c1, c0))));
  }
  publish<2>(acc, 0, 1, partials, counters, slot, seq);
}

// ---- A B at t = 0, 2 ----
template <bool kBind>
__global__ void __launch_bounds__(kRedThreads) k_round_quad(fl_t *A, fl_t *B, size_t q, fl_t r, fl_t *partials, unsigned *counters,
                                                            RoundSlot *slot, uint32_t seq) {
  fl_t acc[2] = {fl_zero(), fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += stride) {
    fl_t a0, a1, b0, b1;
    load_pair<kBind>(A, i, q, r, a0, a1);
    load_pair<kBind>(B, i, q, r, b0, b1);
    acc[0] = fl_add(acc[0], fl_mul(a0, b0));
    fl_t a2 = fl_add(a1, fl_sub(a1, a0)), b2 = fl_add(b1, fl_sub(b1, b0));
    acc[1] = fl_add(acc[1], fl_mul(a2, b2));
  }
  publish<2>(acc, 0, 1, partials, counters, slot, seq);
}

// ---- batched rounds; instance = blockIdx.y ----
// Product-circuit instances (inst < a.nprod) share the third factor eq(rand, .): it is factored out of the round polynomial like
// in k_round_r1cs_split - s(X) = E eq(rand_j, X) t(X) - so the kernel binds A, B and returns t(0) = sum eq_rest A0 B0 and
// t(inf) = sum eq_rest (A1 - A0)(B1 - B0) in vals[3 inst], vals[3 inst + 1] (8 multiplications per thread item instead of 12; the
// eq table is read once per item and never bound). Dot-product instances (own third factor) evaluate A B C at 0, 2, 3 as before.
__device__ __forceinline__ void batched_item_prod(const fl_t &e, const fl_t &a0, const fl_t &a1, const fl_t &b0, const fl_t &b1, fl_t (&acc)[3]) {
  acc[0] = fl_add(acc[0], fl_mul(e, fl_mul(a0, b0)));
  acc[1] = fl_add(acc[1], fl_mul(e, fl_mul(fl_sub(a1, a0), fl_sub(b1, b0))));
}
__device__ __forceinline__ void batched_item_dotp(const fl_t &a0, const fl_t &a1, const fl_t &b0, const fl_t &b1, const fl_t &c0, const fl_t &c1,
                                                  fl_t (&acc)[3]) {
  acc[0] = fl_add(acc[0], fl_mul(fl_mul(a0, b0), c0));
  fl_t da = fl_sub(a1, a0), db = fl_sub(b1, b0), dc = fl_sub(c1, c0);
  fl_t a2 = fl_add(a1, da), b2 = fl_add(b1, db), c2 = fl_add(c1, dc);
  acc[1] = fl_add(acc[1], fl_mul(fl_mul(a2, b2), c2));
  fl_t a3 = fl_add(a2, da), b3 = fl_add(b2, db), c3 = fl_add(c2, dc);
  acc[2] = fl_add(acc[2], fl_mul(fl_mul(a3, b3), c3));
}
template <bool kBind>
__global__ void __launch_bounds__(kRedThreads, 2) k_round_cubic_batched(BatchedRoundArgs a, size_t q, fl_t r, ChalRef ch, fl_t *partials,
                                                                     unsigned *counters, RoundSlot *slot, uint32_t seq) {
  if (kBind && !chal_fetch(ch, r)) return;
  const int inst = blockIdx.y;
  fl_t *A = a.A[inst], *B = a.B[inst];
  fl_t acc[3] = {fl_zero(), fl_zero(), fl_zero()};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if (inst < a.nprod) {
    const fl_t *E = a.eq_rest;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += stride) {
      fl_t a0, a1, b0, b1;
      load_pair<kBind>(A, i, q, r, a0, a1);
      load_pair<kBind>(B, i, q, r, b0, b1);
      batched_item_prod(ldr(E + i), a0, a1, b0, b1, acc);
    }
  } else {
    fl_t *Cc = a.Cout[inst];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += stride) {
      fl_t a0, a1, b0, b1, c0, c1;
      load_pair<kBind>(A, i, q, r, a0, a1);
      load_pair<kBind>(B, i, q, r, b0, b1);
      load_pair<kBind>(Cc, i, q, r, c0, c1);
      batched_item_dotp(a0, a1, b0, b1, c0, c1, acc);
    }
  }
  publish<3>(acc, inst, gridDim.y, partials, counters, slot, seq);
}

// ---- the same for tiny rounds (q <= kSmallQ): ONE block, half a warp per instance, no inter-block traffic at all.
// Roughly 250 of the ~400 rounds of the two product-circuit proofs of a 2^20 instance are of this kind, and their cost
// is pure latency (launch -> loads -> a handful of multiplications -> shuffle tree -> PCIe store).
static const size_t kSmallQ = 64;         // what the kernel supports
static const size_t kSmallQDefault = 16;  // measured on a B200: 16 beats 64 by 0.6 ms per CNN-A proof (a lane then has one item)
template <bool kBind>
__global__ void __launch_bounds__(16 * kMaxBatched + 16) k_round_cubic_batched_small(BatchedRoundArgs a, int ninst, size_t q, fl_t r, ChalRef ch,
                                                                                    RoundSlot *slot, uint32_t seq) {
  if (kBind && !chal_fetch(ch, r)) return;
  const int inst = threadIdx.x >> 4, lane = threadIdx.x & 15;
  fl_t acc[3] = {fl_zero(), fl_zero(), fl_zero()};
  if (inst < ninst) {
    fl_t *A = a.A[inst], *B = a.B[inst];
    if (inst < a.nprod) {
      const fl_t *E = a.eq_rest;
      for (size_t i = lane; i < q; i += 16) {
        fl_t a0, a1, b0, b1;
        load_pair<kBind>(A, i, q, r, a0, a1);
        load_pair<kBind>(B, i, q, r, b0, b1);
        batched_item_prod(ldr(E + i), a0, a1, b0, b1, acc);
      }
    } else {
      fl_t *Cc = a.Cout[inst];
      for (size_t i = lane; i < q; i += 16) {
        fl_t a0, a1, b0, b1, c0, c1;
        load_pair<kBind>(A, i, q, r, a0, a1);
        load_pair<kBind>(B, i, q, r, b0, b1);
        load_pair<kBind>(Cc, i, q, r, c0, c1);
        batched_item_dotp(a0, a1, b0, b1, c0, c1, acc);
      }
    }
  }
  if (q > 1) {  // (q == 1: only lane 0 of each instance holds a term)
#pragma unroll
    for (int off = 8; off > 0; off >>= 1)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        fl_t o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.v[i] = __shfl_down_sync(0xffffffffu, acc[k].v[i], off, 16);
        acc[k] = fl_add(acc[k], o);
      }
  }
  if (lane == 0 && inst < ninst) {
#pragma unroll
    for (int k = 0; k < 3; k++) str(slot->vals + inst * 3 + k, acc[k]);
  }
  // ONE system-scope fence for the block: the barrier orders every lane's stores before thread 0's fence, and a fence is
  // cumulative - what thread 0 has synchronised with is visible before what it writes next (a fence per storing lane as well
  // cost a second PCIe flush on the critical path of every small round)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&slot->seq) = seq;
  }
}

// ---- final claims: vals[k] = p_k[0] + r (p_k[1] - p_k[0])  (or p_k[0] when nothing is left to bind) ----
__global__ void __launch_bounds__(64) k_round_final(FinalArgs a, fl_t r, ChalRef ch, int bind, RoundSlot *slot, uint32_t seq) {
  if (bind && !chal_fetch(ch, r)) return;
  int k = threadIdx.x;
  if (k < a.n) {
    const fl_t *p = a.p[k];
    fl_t v = ldr(p);
    if (bind) v = bind1(v, ldr(p + 1), r);
    str(slot->vals + k, v);
  }
  __syncthreads();  // (one fence for the block, by the thread that publishes: see k_round_cubic_batched_small)
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&slot->seq) = seq;
  }
}

// ---- bullet reduction (Spartan/src/nizk/bullet.rs:72-119), one launch per round ----
// The folded generators are never materialised: W holds, per ORIGINAL generator, the product of the u / u^-1 factors it
// has collected, so L and R are fixed-base MSMs over the original table with scalars a'[partner] * W[j]. This kernel
// (i) applies the previous round's fold to a, b (ping-pong buffers) and W, (ii) accumulates c_L = <a_L, b_R> and
// c_R = <a_R, b_L>, (iii) writes the signed MSM digits of the L row (columns in the upper half of their 2*cur block)
// and the R row (lower half). In `final` mode (vectors of length 1) it emits a'[0], b'[0] and the digits of d * W.
__global__ void __launch_bounds__(kRedThreads) k_bullet_round(BulletRoundArgs p) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t len = p.len, cur = len >> 1;
  const bool fold = p.fold != 0;
  auto A = [&](size_t k) { return fold ? fl_add(fl_mul(ldr(p.a_old + k), p.u), fl_mul(p.uinv, ldr(p.a_old + k + len))) : ldr(p.a_old + k); };
  auto B = [&](size_t k) { return fold ? fl_add(fl_mul(ldr(p.b_old + k), p.uinv), fl_mul(p.u, ldr(p.b_old + k + len))) : ldr(p.b_old + k); };
  fl_t acc[2] = {fl_zero(), fl_zero()};
  if (p.final) {
    if (j == 0) { acc[0] = A(0); acc[1] = B(0); }
  } else if (j < cur) {
    fl_t a_lo = A(j), a_hi = A(j + cur), b_lo = B(j), b_hi = B(j + cur);
    if (fold) { str(p.a_new + j, a_lo); str(p.a_new + j + cur, a_hi); str(p.b_new + j, b_lo); str(p.b_new + j + cur, b_hi); }
    acc[0] = fl_mul(a_lo, b_hi);
    acc[1] = fl_mul(a_hi, b_lo);
  }
  uint32_t nz = 0;
  if (j < p.stride) {
    const size_t plane = (size_t)(p.final ? 1 : 2) * p.stride;
    if (j < p.n) {
      fl_t w = ldr(p.W + j);
      if (fold) {
        w = fl_mul(w, (j & (2 * len - 1)) < len ? p.uinv : p.u);
        str(p.W + j, w);
      }
      if (p.final) {
        nz = msm_recode_value(fl_mul(p.d, w), p.geom, p.digits + j, plane);
      } else {
        size_t rem = j & (len - 1);
        bool upper = rem >= cur;  // a_L against G_R -> row 0 (L); a_R against G_L -> row 1 (R)
        fl_t s = fl_mul(A(upper ? rem - cur : rem + cur), w);
        nz = msm_recode_value(s, p.geom, p.digits + (upper ? 0 : p.stride) + j, plane);
        uint16_t *other = p.digits + (upper ? p.stride : 0) + j;
        for (int wdw = 0; wdw < p.geom.windows; wdw++) other[(size_t)wdw * plane] = 0;
      }
    } else {
      for (int row = 0; row < (p.final ? 1 : 2); row++)
        for (int wdw = 0; wdw < p.geom.windows; wdw++) p.digits[(size_t)wdw * plane + row * p.stride + j] = 0;
    }
  }
  if (p.nonzero) {
    nz = __reduce_add_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0 && nz) atomicAdd(p.nonzero, (unsigned long long)nz);
  }
  publish<2>(acc, 0, 1, p.ctl.d_partials, p.ctl.d_counters, p.ctl.slot, p.ctl.seq);
}
// ---- tail of a batched layer: the first `len` elements of up to kRoundSlotVals tables -> host-mapped memory (the host finishes
// the last rounds of the layer itself, prover.cu batched_prove), then the sequence number ----
__global__ void __launch_bounds__(256) k_tail_copy(FinalArgs a, int len, fl_t *dst, RoundSlot *slot, uint32_t seq) {
  for (int i = threadIdx.x; i < a.n * len; i += blockDim.x) str(dst + i, ldr(a.p[i / len] + (i % len)));
  __syncthreads();  // (one fence for the block, by the thread that publishes: see k_round_cubic_batched_small)
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&slot->seq) = seq;
  }
}
__global__ void k_publish_seq(RoundSlot *slot, uint32_t seq) {
  __threadfence_system();
  *reinterpret_cast<volatile uint32_t *>(&slot->seq) = seq;
}

__global__ void k_stage_vals(const fl_t *src, int n_valid, int n_total, fl_t *dst) {
  int k = threadIdx.x;
  if (k < n_total) str(dst + k, k < n_valid ? ldr(src + k) : fl_zero());
}
__global__ void k_publish_vals(const fl_t *src, int count, RoundSlot *slot, uint32_t seq) {
  int k = threadIdx.x;
  if (k < count) str(slot->vals + k, ldr(src + k));
  __syncthreads();  // (one fence for the block, by the thread that publishes: see k_round_cubic_batched_small)
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&slot->seq) = seq;
  }
}

inline int round_blocks(size_t q, int cap) {
  size_t b = (q + kRedThreads - 1) / kRedThreads;
  if (b < 1) b = 1;
  return (int)(b > (size_t)cap ? cap : b);
}

}  // namespace

void launch_round_cubic_additive(fl_t *A, fl_t *B, fl_t *C, fl_t *D, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st) {
  int nb = round_blocks(q, kRedBlocks);
  ++g_kernel_launches;
  if (bind) k_round_cubic_additive<true><<<nb, kRedThreads, 0, st>>>(A, B, C, D, q, r, c.d_partials, c.d_counters, c.slot, c.seq);
  else k_round_cubic_additive<false><<<nb, kRedThreads, 0, st>>>(A, B, C, D, q, r, c.d_partials, c.d_counters, c.slot, c.seq);
}
void launch_round_r1cs_split(const fl_t *eq_rest, fl_t *B, fl_t *C, fl_t *D, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st) {
  int nb = round_blocks(q, kRedBlocks);
  ++g_kernel_launches;
  if (bind) k_round_r1cs_split<true><<<nb, kRedThreads, 0, st>>>(eq_rest, B, C, D, q, r, c.d_partials, c.d_counters, c.slot, c.seq);
  else k_round_r1cs_split<false><<<nb, kRedThreads, 0, st>>>(eq_rest, B, C, D, q, r, c.d_partials, c.d_counters, c.slot, c.seq);
}
void launch_round_quad(fl_t *A, fl_t *B, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st) {
  int nb = round_blocks(q, kRedBlocks);
  ++g_kernel_launches;
  if (bind) k_round_quad<true><<<nb, kRedThreads, 0, st>>>(A, B, q, r, c.d_partials, c.d_counters, c.slot, c.seq);
  else k_round_quad<false><<<nb, kRedThreads, 0, st>>>(A, B, q, r, c.d_partials, c.d_counters, c.slot, c.seq);
}
static size_t small_q_threshold() {  // VPIN_SMALL_Q overrides (experiments)
  static const size_t v = [] {
    const char *e = getenv("VPIN_SMALL_Q");
    long x = e ? atol(e) : (long)kSmallQDefault;
    return (size_t)(x < 0 ? 0 : (x > (long)kSmallQ ? (long)kSmallQ : x));
  }();
  return v;
}
void launch_round_cubic_batched(const BatchedRoundArgs &a, int ninst, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st,
                                const ChalRef &ch) {
  ++g_kernel_launches;
  if (q <= small_q_threshold()) {
    int threads = (16 * ninst + 31) / 32 * 32;
    if (bind) k_round_cubic_batched_small<true><<<1, threads, 0, st>>>(a, ninst, q, r, ch, c.slot, c.seq);
    else k_round_cubic_batched_small<false><<<1, threads, 0, st>>>(a, ninst, q, r, ch, c.slot, c.seq);
    return;
  }
  int nb = round_blocks(q, ninst > 4 ? kRedBlocks / 4 : kRedBlocks);
  dim3 grid(nb, ninst);
  if (bind) k_round_cubic_batched<true><<<grid, kRedThreads, 0, st>>>(a, q, r, ch, c.d_partials, c.d_counters, c.slot, c.seq);
  else k_round_cubic_batched<false><<<grid, kRedThreads, 0, st>>>(a, q, r, ch, c.d_partials, c.d_counters, c.slot, c.seq);
}
void launch_bullet_round(const BulletRoundArgs &p, cudaStream_t st) {
  unsigned nb = (unsigned)((p.stride + kRedThreads - 1) / kRedThreads);
  ++g_kernel_launches, k_bullet_round<<<nb, kRedThreads, 0, st>>>(p);
}
void launch_stage_vals(const fl_t *src, int n_valid, int n_total, fl_t *dst, cudaStream_t st) {
  ++g_kernel_launches, k_stage_vals<<<1, 128, 0, st>>>(src, n_valid, n_total, dst);
}
void launch_publish_vals(const fl_t *src, int count, RoundSlot *slot, uint32_t seq, cudaStream_t st) {
  ++g_kernel_launches, k_publish_vals<<<1, 128, 0, st>>>(src, count, slot, seq);
}
void launch_publish_seq(RoundSlot *slot, uint32_t seq, cudaStream_t st) { ++g_kernel_launches, k_publish_seq<<<1, 1, 0, st>>>(slot, seq); }
void launch_tail_copy(const FinalArgs &a, int len, fl_t *d_dst, const RoundCtl &c, cudaStream_t st) {
  ++g_kernel_launches, k_tail_copy<<<1, 256, 0, st>>>(a, len, d_dst, c.slot, c.seq);
}
void launch_round_final(const FinalArgs &a, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st, const ChalRef &ch) {
  ++g_kernel_launches, k_round_final<<<1, 64, 0, st>>>(a, r, ch, bind ? 1 : 0, c.slot, c.seq);
}

}  // namespace vpin
