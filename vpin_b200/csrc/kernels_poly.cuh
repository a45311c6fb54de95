// Launch wrappers of the F_l table kernels (sumcheck rounds, eq tables, binds, SpMV, SPARK layers).
// All tables are arrays of fl_t in Montgomery form in HBM. Every wrapper enqueues on `st` and returns.
#pragma once
#include <cuda_runtime.h>

#include "fl.cuh"
#include "kernels_msm.cuh"

namespace vpin {

// number of partial-sum slots the reduction kernels need (fl_t elements): kMaxBlocks * k per instance
static const int kRedBlocks = 592;  // 148 SMs x 4
static const int kRedThreads = 256;

// K4: eq(r) table, r[0] = most significant index bit (Spartan/src/dense_mlpoly.rs:78-94).
// d_r: ell challenges on device. d_out: 2^ell. d_tmp: >= max(4096, 3 * 2^ceil(ell/2)) scratch elements.
void launch_eq_evals(const fl_t *d_r, int ell, fl_t *d_out, fl_t *d_tmp, cudaStream_t st);

// K5 bind: Z[i] += r * (Z[i + half] - Z[i]) for i < half (Spartan/src/dense_mlpoly.rs:229-236); r on device.
void launch_bind_top(fl_t *Z, size_t half, const fl_t *d_r, cudaStream_t st);
// several equally long tables in one launch (ptrs on device)
void launch_bind_top_multi(fl_t *const *d_tables, int ntables, size_t half, const fl_t *d_r, cudaStream_t st);
// Z[i] = Z[2i] + r (Z[2i+1] - Z[2i]) into out (Spartan/src/dense_mlpoly.rs:238-245)
void launch_bind_bot(const fl_t *Z, fl_t *out, size_t half, const fl_t *d_r, cudaStream_t st);

// K5 eval: sum over i < half of comb at t = 0, 2, 3 with comb = A (B C - D) (Spartan/src/sumcheck.rs:619-652).
// d_out: 3 elements; d_partials: 3 * kRedBlocks scratch.
void launch_cubic_additive_round(const fl_t *A, const fl_t *B, const fl_t *C, const fl_t *D, size_t half, fl_t *d_out,
                                 fl_t *d_partials, cudaStream_t st);
// K6 eval: comb = A B at t = 0, 2 (Spartan/src/sumcheck.rs:456-469). d_out: 2 elements.
void launch_quad_round(const fl_t *A, const fl_t *B, size_t half, fl_t *d_out, fl_t *d_partials, cudaStream_t st);
// K7 eval: for each instance k < n: comb = A_k B_k C_k at t = 0, 2, 3 (Spartan/src/sumcheck.rs:287-357).
// d_A/d_B/d_C: device arrays of n table pointers (C may repeat the shared eq table). d_out: 3n elements.
// d_partials: 3 * kRedBlocks * n scratch.
void launch_cubic_batched_round(const fl_t *const *d_A, const fl_t *const *d_B, const fl_t *const *d_C, int n, size_t half,
                                fl_t *d_out, fl_t *d_partials, cudaStream_t st);

// ---- fused rounds (kernels_round.cu): bind with the previous challenge + evaluate the next round in one launch, results
// delivered to host-mapped pinned memory. ----
static const int kRoundSlotVals = 96;  // >= 3 x the instances of a batched layer spread over up to 8 ranks (sharded rounds)
struct RoundSlot {                 // lives in cudaHostAllocMapped memory
  fl_t vals[kRoundSlotVals];
  uint32_t seq;                    // written last (after a system-wide fence) with the launch's sequence number
  uint32_t pad[7];
};
struct RoundCtl {
  fl_t *d_partials;                // >= 3 * kRedBlocks * ninst elements
  unsigned *d_counters;            // 1 + ninst words, zero between launches
  RoundSlot *slot;                 // device-visible address of the slot
  uint32_t seq;
};
// Challenge mailbox: a round kernel can be launched BEFORE the challenge it binds with is known - it then waits for the host to
// post the challenge in host-mapped memory instead of receiving it as a kernel parameter. The launch latency (API call + front
// end, ~6 us of a ~15 us round trip) leaves the critical path of the latency-bound rounds; the Fiat-Shamir transcript stays on
// the host. A slot is eight 8-byte atoms (limb, tag): each atom validates itself, so one 64-byte read is a complete poll no
// matter in which order the host's stores land. In a multi-block grid the first block to arrive polls the host and relays the
// value through a device-side latch (one PCIe poller per kernel). A poll that outlives ChalRef::timeout_ms (the host died, or
// was held up for seconds) sets the latch's sticky abort word: that kernel and every later mailbox kernel of the context
// returns without publishing, which the host's round_wait reports as an error.
struct ChalSlot { uint32_t w[16]; };                                   // host-mapped
struct ChalLatch { uint32_t owner, ready, abort, pad[5]; uint32_t r[8]; };  // device memory, one per context, zero at creation
struct ChalRef { const ChalSlot *slot; ChalLatch *latch; uint32_t tag, timeout_ms; };  // slot == nullptr: the challenge is the kernel parameter
static const int kChalRing = 64;
static const int kMaxBatched = 18;  // 12 product circuits + 6 dot-product halves (Spartan/src/sparse_mlpoly.rs:1173-1197)
struct BatchedRoundArgs {
  fl_t *A[kMaxBatched], *B[kMaxBatched];  // bound in place
  const fl_t *Cin[kMaxBatched];           // third factor of a dot-product instance (its own table) ...
  fl_t *Cout[kMaxBatched];                // ... bound in place (Cout == Cin); unused for product instances
  int nprod;                              // instances [0, nprod) are product circuits: their third factor is the shared eq
  const fl_t *eq_rest;                    // polynomial, factored out of the round polynomial: eq(rand[j+1..], .) of this
};                                        // round (q elements, read only); see k_round_cubic_batched
struct FinalArgs { const fl_t *p[kRoundSlotVals]; int n; };
// q = number of thread items: half the current length without bind, a quarter of it with bind (the tables then shrink
// to half their length in place). r is ignored without bind.
void launch_round_cubic_additive(fl_t *A, fl_t *B, fl_t *C, fl_t *D, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st);
// the same round with the eq factor A = eq(tau, .) factored out (s(X) = eq(tau_j, X) E t(X), t quadratic): binds B, C, D and
// delivers t(0) = sum eq_rest (B C - D) at X = 0 and the leading coefficient t(inf) = sum eq_rest (B1 - B0)(C1 - C0)
void launch_round_r1cs_split(const fl_t *eq_rest, fl_t *B, fl_t *C, fl_t *D, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st);
void launch_round_quad(fl_t *A, fl_t *B, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st);
void launch_round_cubic_batched(const BatchedRoundArgs &a, int ninst, size_t q, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st,
                                const ChalRef &ch = ChalRef{nullptr, nullptr, 0, 0});
// dst[t * len + i] = p_t[i] for t < a.n, i < len (dst: device address of host-mapped memory), then c.slot->seq = c.seq
static const int kTailElems = 1024;  // capacity of the mapped tail buffer (elements)
void launch_tail_copy(const FinalArgs &a, int len, fl_t *d_dst, const RoundCtl &c, cudaStream_t st);
// vals[k] = p_k[0] + r (p_k[1] - p_k[0]) (bind) or p_k[0]
void launch_round_final(const FinalArgs &a, bool bind, const fl_t &r, const RoundCtl &c, cudaStream_t st,
                        const ChalRef &ch = ChalRef{nullptr, nullptr, 0, 0});
// One bullet-reduction round (Spartan/src/nizk/bullet.rs:72-119) in the fixed-base formulation, see kernels_round.cu.
struct BulletRoundArgs {
  const fl_t *a_old, *b_old;  // vectors before the pending fold (length 2 * len when fold, len otherwise)
  fl_t *a_new, *b_new;        // folded vectors, length len (written when fold and not final)
  fl_t *W;                    // n weights over the original generators, updated in place when fold
  size_t n, len;              // len: vector length after the pending fold
  fl_t u, uinv, d;            // previous challenge; d: the final blind multiplier (final mode)
  int fold, final;
  MsmGeom geom;               // window geometry of the generator table the rows are multiplied against
  uint16_t *digits;           // MSM digits [window][row][stride], rows = final ? 1 : 2 (row 0 = L, row 1 = R)
  size_t stride;
  unsigned long long *nonzero;
  RoundCtl ctl;               // vals[0], vals[1] = c_L, c_R (or a'[0], b'[0] in final mode)
};
void launch_bullet_round(const BulletRoundArgs &p, cudaStream_t st);
void launch_publish_seq(RoundSlot *slot, uint32_t seq, cudaStream_t st);
// sharded rounds (one proof on several GPUs): dst[0..n_total) = src[0..n_valid) followed by zeros (this rank's segment of the
// all-gather buffer); then, after the all-gather, slot->vals[0..count) = src[0..count) and the sequence number
void launch_stage_vals(const fl_t *src, int n_valid, int n_total, fl_t *dst, cudaStream_t st);
void launch_publish_vals(const fl_t *src, int count, RoundSlot *slot, uint32_t seq, cudaStream_t st);
// eq table with the point passed by value (no device-side copy of r needed); same output as launch_eq_evals
struct EqPoint { fl_t r[32]; };
void launch_eq_evals_pt(const EqPoint &pt, int ell, fl_t *d_out, fl_t *d_tmp, cudaStream_t st);
// suffix tables of eq(pt, .): S[2^k + x] = eq(pt.r[ell-k .. ell), x) for k = 0..kmax (kmax <= ell), x < 2^k; S has 2^(kmax+1)
// elements (S[0] unused). Round j of a sumcheck over eq(pt, .) g(.) reads table k = ell - 1 - j (kernels_round.cu).
void launch_eq_suffix(const EqPoint &pt, int ell, int kmax, fl_t *S, cudaStream_t st);

// dot product sum_i A[i] B[i] (Spartan/src/nizk/mod.rs:442-445). d_out: 1 element; d_partials: kRedBlocks.
void launch_dot(const fl_t *A, const fl_t *B, size_t n, fl_t *d_out, fl_t *d_partials, cudaStream_t st);
// n dot products against one shared vector B: out[k] = <A_k, B>
void launch_dot_multi(const fl_t *const *d_A, const fl_t *B, int n, size_t len, fl_t *d_out, fl_t *d_partials, cudaStream_t st);
// sum_i A[i] B[i] C[i]
void launch_dot3(const fl_t *A, const fl_t *B, const fl_t *C, size_t n, fl_t *d_out, fl_t *d_partials, cudaStream_t st);

// K10: LZ[j] = sum_i L[i] Z[i * R + j] (Spartan/src/dense_mlpoly.rs:220-227). d_tmp: nsplit * R scratch (nsplit <= 64).
void launch_bound(const fl_t *Z, const fl_t *L, size_t Lsize, size_t Rsize, fl_t *d_out, fl_t *d_tmp, cudaStream_t st);

// Value dictionary of the sparse matrices: code of an entry's coefficient. vPIN's constraint systems use +-1 for 93 % of the
// entries and +-2, 3 for almost all others; those are additions, not multiplications, and their 32-byte value is never read.
enum SpmvCode : uint8_t { kCodeGeneral = 0, kCodePlus1 = 1, kCodeMinus1 = 2, kCodePlus2 = 3, kCodeMinus2 = 4, kCodePlus3 = 5 };
// K2: CSR SpMV out[row] = sum val * z[col] (Spartan/src/sparse_mlpoly.rs:467-481). One thread per row for the short rows
// (1-3 entries: consecutive threads read consecutive entries, so code / index loads coalesce), one warp per long row
// (lanes stride the entries, shuffle-tree reduction): a 128-term bit-decomposition row no longer serialises in one thread.
static const int kLongRow = 16;
struct CsrDev { const uint32_t *ptr; const uint32_t *idx; const fl_t *val; size_t n; const uint8_t *code; const uint32_t *long_rows; size_t n_long_rows; };
void launch_spmv_csr(const CsrDev &m, const fl_t *z, fl_t *out, cudaStream_t st);
// K3: CSC form of M^T x (Spartan/src/sparse_mlpoly.rs:483-498): out[col] = sum x[row] * val, accumulated with a scale:
// out[col] (+)= scale * sum. Columns with more than kLongCol entries are listed in long_cols and reduced by a block each.
static const int kLongCol = 256;
struct CscDev { const uint32_t *ptr; const uint32_t *idx; const fl_t *val; size_t n; const uint32_t *long_cols; size_t n_long; const uint8_t *code; };
// d_scratch: scratch_elems >= n_long elements (up to 64 slices per long column are used)
void launch_spmv_csc_scaled(const CscDev &m, const fl_t *x, const fl_t *d_scale, bool accumulate, fl_t *out, fl_t *d_scratch,
                            size_t scratch_elems, cudaStream_t st);
// sum over nnz of rx[row] * ry[col] * val (Spartan/src/sparse_mlpoly.rs:440-452); uses the CSR arrays + a row index per entry
void launch_sparse_eval(const uint32_t *rows, const uint32_t *cols, const fl_t *val, size_t nnz, const fl_t *trx, const fl_t *try_,
                        fl_t *d_out, fl_t *d_partials, cudaStream_t st);

// SPARK (Spartan/src/sparse_mlpoly.rs)
// deref gather out[i] = mem[addr[i]]  (:267-276)
void launch_gather(const uint32_t *addr, const fl_t *mem, size_t n, fl_t *out, cudaStream_t st);
// out[i] = fl(u32 in[i])
void launch_u32_to_fl(const uint32_t *in, size_t n, fl_t *out, cudaStream_t st);
// in-place exclusive prefix sum of n u32 counters (multi-block); d_scratch: exclusive_scan_scratch_words(n) words
size_t exclusive_scan_scratch_words(size_t n);
void launch_exclusive_scan_u32(uint32_t *d, size_t n, uint32_t *d_scratch, cudaStream_t st);
// memory-checking timestamps (:232-265) by a stable radix sort instead of the reference's sequential replay (kernels_sort.cu).
// addr[k]: nnz[k] addresses of matrix k (row or column indices, COO order), padded with address 0 to N operations each.
// Outputs: d_addr_out, d_read_ts (3N words, A | B | C) and d_audit_ts (M words). d_scratch: spark_timestamps_scratch_words.
size_t spark_timestamps_scratch_words(size_t N, size_t M);
void launch_spark_timestamps(const uint32_t *const addr[3], const size_t nnz[3], size_t N, size_t M, uint32_t *d_addr_out, uint32_t *d_read_ts,
                             uint32_t *d_audit_ts, uint32_t *d_scratch, cudaStream_t st);
// hash layer (:547-622): h = ts*gamma^2 + val*gamma + addr - tau. d_gt: {gamma, tau} on device.
//   init[i]  = eq[i]*gamma + i - tau ; audit[i] = audit_ts[i]*gamma^2 + eq[i]*gamma + i - tau      (i < num_cells)
void launch_hash_mem(const fl_t *eq, const uint32_t *audit_ts, size_t num_cells, const fl_t *d_gt, fl_t *init, fl_t *audit, cudaStream_t st);
//   read[i]  = ts[i]*gamma^2 + deref[i]*gamma + addr[i] - tau ; write[i] = read[i] + gamma^2       (i < num_ops)
void launch_hash_ops(const uint32_t *addr, const fl_t *deref, const uint32_t *read_ts, size_t num_ops, const fl_t *d_gt, fl_t *read,
                     fl_t *write, cudaStream_t st);
// product tree layer (Spartan/src/product_tree.rs:18-34): out[i] = in[i] * in[i + n] for i < n (out may not alias in)
void launch_mul_halves(const fl_t *in, size_t n, fl_t *out, cudaStream_t st);
// packed product trees (layers n, n/2, ..., 2 of each tree, leaves already at p[k][0..n)) of up to 16 circuits at once: one
// launch per layer for all trees while a layer is longer than kTreeTail, then one block per tree for the rest
static const size_t kTreeTail = 8192;
struct TreeBatch { fl_t *p[16]; int n; };
void launch_build_trees(const TreeBatch &b, size_t n, cudaStream_t st);
// out[i] = a*A[i] + b*B[i] + c*C[i] ; d_abc = {a,b,c} on device
void launch_lincomb3(const fl_t *A, const fl_t *B, const fl_t *C, const fl_t *d_abc, size_t n, fl_t *out, cudaStream_t st);

// bullet reduction helpers (Spartan/src/nizk/bullet.rs:72-119), see prover for the fixed-base reformulation.
// a[i] = a[i]*u + uinv*a[i+n]; b[i] = b[i]*uinv + u*b[i+n]  (i < n). d_u = {u, uinv}
void launch_bullet_fold(fl_t *a, fl_t *b, size_t n, const fl_t *d_u, cudaStream_t st);
// w[j] *= ((j mod 2n) < n ? uinv : u) for j < total
void launch_bullet_weights(fl_t *w, size_t total, size_t n, const fl_t *d_u, cudaStream_t st);
// scalar rows for the L and R multiscalar multiplications over the ORIGINAL generators:
// sL[j] = (j mod 2n) >= n ? a[(j mod 2n) - n] * w[j] : 0 ;  sR[j] = (j mod 2n) < n ? a[(j mod 2n) + n] * w[j] : 0
void launch_bullet_scalars(const fl_t *a, const fl_t *w, size_t total, size_t n, fl_t *sL, fl_t *sR, cudaStream_t st);
// out[i] = s * in[i]; d_s on device
void launch_scale(const fl_t *in, const fl_t *d_s, size_t n, fl_t *out, cudaStream_t st);
// out[i] = a[i] + b[i]
void launch_add_vec(const fl_t *a, const fl_t *b, size_t n, fl_t *out, cudaStream_t st);
void launch_fill_one(fl_t *out, size_t n, cudaStream_t st);
// Montgomery <-> canonical conversion of bulk arrays (C ABI takes canonical little-endian scalars)
// canonical -> Montgomery with the range check of Scalar::from_bytes; *d_bad |= 1 if any input >= l (in == out allowed)
void launch_from_bytes_checked(const fl_t *in, size_t n, fl_t *out, uint32_t *d_bad, cudaStream_t st);
void launch_to_mont(const fl_t *in, size_t n, fl_t *out, cudaStream_t st);
void launch_from_mont(const fl_t *in, size_t n, fl_t *out, cudaStream_t st);

}  // namespace vpin
