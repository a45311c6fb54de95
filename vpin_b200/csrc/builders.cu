// vPIN's R1CS gadgets as fixture generators for the B200 prover:
//   vPIN_proof_generation/src/point_addition.rs:5-326 — affine EC addition P + R with an infinity flag, 10 constraints
//   vPIN_proof_generation/src/point_mult.rs:7-704     — 128-step double-and-add, 27n+8 constraints per multiplication
// The constraint systems are emitted from small per-constraint tables (the order of entries inside each of A, B, C is
// the reference's, because SPARK commits to the COO order). Witnesses are expanded step-synchronously across all
// multiplications so the two field inversions per step are batched (Montgomery's trick).
#include "prover.cuh"

namespace vpin {

namespace {
enum Coef { ONE, MINUS_ONE, TWO, THREE, MINUS_TWO };
struct Term { long col; Coef k; };  // col >= 0: variable offset inside the gadget block; col == -1: constant 1; -2: public input 0

struct Emitter {
  std::vector<vpin_coo_entry> M[3];
  uint8_t coef[5][32];
  size_t num_vars = 0;
  Emitter() {
    fl_t one = fl_one(), two = one + one;
    fl_to_bytes(one, coef[ONE]);
    fl_to_bytes(fl_neg(one), coef[MINUS_ONE]);
    fl_to_bytes(two, coef[TWO]);
    fl_to_bytes(two + one, coef[THREE]);
    fl_to_bytes(fl_neg(two), coef[MINUS_TWO]);
  }
  void term(int mat, size_t row, size_t block, const Term &t) {
    vpin_coo_entry e;
    e.row = row;
    e.col = t.col >= 0 ? block + (size_t)t.col : (t.col == -1 ? num_vars : num_vars + 1);
    memcpy(e.val, coef[t.k], 32);
    M[mat].push_back(e);
  }
  void con(size_t row, size_t block, std::initializer_list<Term> a, std::initializer_list<Term> b, std::initializer_list<Term> c) {
    for (auto &t : a) term(0, row, block, t);
    for (auto &t : b) term(1, row, block, t);
    for (auto &t : c) term(2, row, block, t);
  }
};
const long K1 = -1, IN0 = -2;

void batch_invert(std::vector<fl_t> &v) {  // zeros stay zero (dalek's invert maps 0 to 0)
  std::vector<fl_t> prefix(v.size());
  fl_t run = fl_one();
  for (size_t i = 0; i < v.size(); i++) {
    prefix[i] = run;
    if (!fl_is_zero(v[i])) run = run * v[i];
  }
  fl_t inv = fl_invert(run);
  for (size_t i = v.size(); i-- > 0;) {
    if (fl_is_zero(v[i])) continue;
    fl_t x = inv * prefix[i];
    inv = inv * v[i];
    v[i] = x;
  }
}
fl_t from_bytes_mod_order(const uint8_t *b) {  // dalek Scalar::from_bytes_mod_order
  uint8_t w[64];
  memcpy(w, b, 32);
  memset(w + 32, 0, 32);
  return fl_from_bytes_wide(w);
}
}  // namespace

// point_addition.rs:38-70
void point_add_dims(uint64_t n, uint64_t d[4]) {
  uint64_t p1, p2, p3;
  if (n < 780) { p1 = 2; p2 = 25; p3 = 3; }
  else if (n > 2130 && n < 2150) { p1 = 5; p2 = 30; p3 = 5; }
  else if (n > 2149 && n < 2450) { p1 = 3; p2 = 30; p3 = 5; }
  else if (n > 5000 && n < 8000) { p1 = 3; p2 = 20; p3 = 5; }
  else { p1 = 5; p2 = 30; p3 = 5; }
  d[0] = 10 * n; d[1] = 15 * n + 1; d[2] = 0; d[3] = p1 * (p2 / p3) * n;
}
// point_mult.rs:27-67 (n = 128 bits, load_data.rs:62)
void point_mult_dims(uint64_t m, uint64_t d[4]) {
  const uint64_t n = 128;
  uint64_t p1, p2, p3;
  if (m == 50) { p1 = 100; p2 = 2; p3 = 80; }
  else if (m == 210) { p1 = 300; p2 = 2; p3 = 20; }
  else if (m == 240) { p1 = 300; p2 = 4; p3 = 20; }
  else if (m < 660) { p1 = 100; p2 = 2; p3 = 40; }
  else if (m == 6000) { p1 = 250; p2 = 2; p3 = 20; }
  else { p1 = 350; p2 = 2; p3 = 20; }
  d[0] = (27 * n + 8) * m; d[1] = (27 * n + 10) * m + 1; d[2] = 1; d[3] = p1 * (p2 * n + p3 * m);
}

std::unique_ptr<Instance> build_point_add(Ctx *ctx, uint64_t n, const uint8_t *px32, const uint8_t *py32, const uint8_t *rx32,
                                          const uint8_t *ry32, const int64_t *rz_flags, uint64_t dims[4], uint8_t *vars_para32,
                                          uint8_t *vars_input32, uint8_t *vars32) {
  point_add_dims(n, dims);
  Emitter E;
  E.num_vars = dims[1];
  // block of 15: c, Rx, Px, Ry, Py, Rz, s1, s2, s3, t1, t2, t3, t4, x3, y3   (point_addition.rs:68)
  enum { c = 0, Rx, Px, Ry, Py, Rz, s1, s2, s3, t1, t2, t3, t4, x3, y3 };
  for (size_t i = 0; i < n; i++) {  // :81-151
    size_t r = 10 * i, b = 15 * i;
    E.con(r + 0, b, {{c, ONE}}, {{Rx, ONE}, {Px, MINUS_ONE}}, {{K1, ONE}});
    E.con(r + 1, b, {{Ry, ONE}, {Py, MINUS_ONE}}, {{c, ONE}}, {{s1, ONE}});
    E.con(r + 2, b, {{s1, ONE}}, {{s1, ONE}}, {{s2, ONE}});
    E.con(r + 3, b, {{s2, ONE}, {Px, MINUS_ONE}, {Rx, MINUS_ONE}}, {{K1, ONE}, {Rz, MINUS_ONE}}, {{t1, ONE}});
    E.con(r + 4, b, {{Px, ONE}}, {{Rz, ONE}}, {{t2, ONE}});
    E.con(r + 5, b, {{t1, ONE}, {t2, ONE}}, {{K1, ONE}}, {{x3, ONE}});
    E.con(r + 6, b, {{s1, ONE}}, {{Px, ONE}, {x3, MINUS_ONE}}, {{s3, ONE}});
    E.con(r + 7, b, {{s3, ONE}, {Py, MINUS_ONE}}, {{K1, ONE}, {Rz, MINUS_ONE}}, {{t3, ONE}});
    E.con(r + 8, b, {{Py, ONE}}, {{Rz, ONE}}, {{t4, ONE}});
    E.con(r + 9, b, {{t3, ONE}, {t4, ONE}}, {{K1, ONE}}, {{y3, ONE}});
  }
  // witness :157-267
  size_t nv = dims[1];
  memset(vars_para32, 0, 32 * nv);
  memset(vars_input32, 0, 32 * nv);
  std::vector<fl_t> inv(n), P_x(n), P_y(n), R_x(n), R_y(n);
  for (size_t i = 0; i < n; i++) {
    P_x[i] = from_bytes_mod_order(px32 + 32 * i); P_y[i] = from_bytes_mod_order(py32 + 32 * i);
    R_x[i] = from_bytes_mod_order(rx32 + 32 * i); R_y[i] = from_bytes_mod_order(ry32 + 32 * i);
    inv[i] = R_x[i] - P_x[i];
  }
  batch_invert(inv);
  fl_t one = fl_one();
  for (size_t i = 0; i < n; i++) {
    fl_t rz = rz_flags[i] == 0 ? fl_zero() : one;
    fl_t vc = inv[i];
    fl_t vs1 = (R_y[i] - P_y[i]) * vc;
    fl_t vs2 = vs1 * vs1;
    fl_t vt1 = (vs2 - P_x[i] - R_x[i]) * (one - rz);
    fl_t vt2 = P_x[i] * rz;
    fl_t vx3 = vt1 + vt2;
    fl_t vs3 = vs1 * (P_x[i] - vx3);
    fl_t vt3 = (vs3 - P_y[i]) * (one - rz);
    fl_t vt4 = P_y[i] * rz;
    fl_t vy3 = vt3 + vt4;
    fl_t blockv[15] = {vc, R_x[i], P_x[i], R_y[i], P_y[i], rz, vs1, vs2, vs3, vt1, vt2, vt3, vt4, vx3, vy3};
    for (int k = 0; k < 15; k++) fl_to_bytes(blockv[k], vars_input32 + 32 * (15 * i + k));
  }
  memcpy(vars32, vars_input32, 32 * nv);
  return instance_create(ctx, dims[0], dims[1], dims[2], E.M[0].data(), E.M[0].size(), E.M[1].data(), E.M[1].size(), E.M[2].data(),
                         E.M[2].size());
}

// ------------------------------------------------------------------------------------------------ point_mult on the device
// point_mult.rs:7-704 is inside the reference's timed region (proof_point_mult.rs:24-101): 256 field inversions per
// multiplication and 12.9 K COO triples per multiplication on one core — 8.4 s for LeNet layer 5 with the host expansion
// that used to live here, eight times the proof itself. Both halves now run on the device:
//  * the constraint system is a closed-form pattern: the (row, col, value) triples of ONE multiplication block are emitted
//    on the host by the table below (entry order inside A, B, C is the reference's, because SPARK commits to the COO order)
//    and k_pm_emit replicates them for every multiplication with the block's row / variable offsets;
//  * the witness is expanded by the k_pm_chain / k_pm_invert / k_pm_affine / k_pm_fill pipeline below (Jacobian chains, two
//    batched inversions per multiplication instead of 256 sequential ones) and written in Montgomery form where the prover
//    reads it.
namespace {
const uint64_t kColSentinel = (uint64_t)1 << 40;  // template column codes: sentinel = constant 1, sentinel + 1 = public input 0
const long PM_N = 128;                            // bits per scalar (load_data.rs:62)
// variable layout of one block (point_mult.rs:517-601)
enum : long {
  BIT = 0, A_SC = PM_N, AX = PM_N + 1, AY = 2 * PM_N + 2, BX = 3 * PM_N + 3, BY = 4 * PM_N + 4, BZ = 5 * PM_N + 5, CX = 6 * PM_N + 6,
  CY = 7 * PM_N + 6, DX = 8 * PM_N + 6, DY = 9 * PM_N + 6, QX = 10 * PM_N + 6, QY = 10 * PM_N + 7, PX = 10 * PM_N + 8, PY = 10 * PM_N + 9,
  C_PA = 10 * PM_N + 10, S1_PA = 11 * PM_N + 10, S2_PA = 12 * PM_N + 10, S3_PA = 13 * PM_N + 10, T1_PA = 14 * PM_N + 10,
  T2_PA = 15 * PM_N + 10, T3_PA = 16 * PM_N + 10, T4_PA = 17 * PM_N + 10, C_PD = 18 * PM_N + 10, T1_PD = 19 * PM_N + 10,
  S1_PD = 20 * PM_N + 10, S2_PD = 21 * PM_N + 10, T2_PD = 22 * PM_N + 10, Z1 = 23 * PM_N + 10, Z2 = 24 * PM_N + 10, Z3 = 25 * PM_N + 10,
  Z4 = 26 * PM_N + 10
};
const size_t PM_ONC = 27 * PM_N + 8, PM_ONV = 27 * PM_N + 10;

// the constraints of one multiplication block with row offset R and variable offset V (point_mult.rs:85-322)
void emit_point_mult_block(Emitter &E, size_t R, size_t V) {
  const long n = PM_N;
  uint8_t pow2[32];
  fl_t tb = fl_one(), two = fl_one() + fl_one();
  for (long i = 0; i < n; i++) {  // sum_i 2^i bit_i * 1 = a
    vpin_coo_entry e;
    fl_to_bytes(tb, pow2);
    tb = tb * two;
    e.row = R; e.col = V + i;
    memcpy(e.val, pow2, 32);
    E.M[0].push_back(e);
  }
  E.con(R, V, {}, {{K1, ONE}}, {{A_SC, ONE}});
  for (long i = 1; i <= n; i++) E.con(R + i, V, {{BIT + i - 1, ONE}}, {{BIT + i - 1, ONE}}, {{BIT + i - 1, ONE}});  // booleanity
  E.con(R + n + 1, V, {{AX, ONE}, {PX, MINUS_ONE}}, {{K1, ONE}}, {});
  E.con(R + n + 2, V, {{AY, ONE}, {PY, MINUS_ONE}}, {{K1, ONE}}, {});
  E.con(R + n + 3, V, {{BX, ONE}}, {{K1, ONE}}, {});
  E.con(R + n + 4, V, {{BY, ONE}}, {{K1, ONE}}, {});
  E.con(R + n + 5, V, {{BZ, ONE}, {K1, MINUS_ONE}}, {{K1, ONE}}, {});
  for (long i = 0; i < n; i++) {
    size_t r = R + n + 26 * i;
    // C = B + A (point addition with infinity flag Bz)  :129-198
    E.con(r + 6, V, {{C_PA + i, ONE}}, {{BX + i, ONE}, {AX + i, MINUS_ONE}}, {{K1, ONE}});
    E.con(r + 7, V, {{BY + i, ONE}, {AY + i, MINUS_ONE}}, {{C_PA + i, ONE}}, {{S1_PA + i, ONE}});
    E.con(r + 8, V, {{S1_PA + i, ONE}}, {{S1_PA + i, ONE}}, {{S2_PA + i, ONE}});
    E.con(r + 9, V, {{S2_PA + i, ONE}, {AX + i, MINUS_ONE}, {BX + i, MINUS_ONE}}, {{K1, ONE}, {BZ + i, MINUS_ONE}}, {{T1_PA + i, ONE}});
    E.con(r + 10, V, {{AX + i, ONE}}, {{BZ + i, ONE}}, {{T2_PA + i, ONE}});
    E.con(r + 11, V, {{T1_PA + i, ONE}, {T2_PA + i, ONE}}, {{K1, ONE}}, {{CX + i, ONE}});
    E.con(r + 12, V, {{S1_PA + i, ONE}}, {{AX + i, ONE}, {CX + i, MINUS_ONE}}, {{S3_PA + i, ONE}});
    E.con(r + 13, V, {{S3_PA + i, ONE}, {AY + i, MINUS_ONE}}, {{K1, ONE}, {BZ + i, MINUS_ONE}}, {{T3_PA + i, ONE}});
    E.con(r + 14, V, {{AY + i, ONE}}, {{BZ + i, ONE}}, {{T4_PA + i, ONE}});
    E.con(r + 15, V, {{T3_PA + i, ONE}, {T4_PA + i, ONE}}, {{K1, ONE}}, {{CY + i, ONE}});
    // D = 2A (point doubling, curve coefficient a is public input 0)  :206-250
    E.con(r + 16, V, {{C_PD + i, ONE}}, {{AY + i, TWO}}, {{K1, ONE}});
    E.con(r + 17, V, {{AX + i, ONE}}, {{AX + i, ONE}}, {{T1_PD + i, ONE}});
    E.con(r + 18, V, {{T1_PD + i, THREE}, {IN0, ONE}}, {{C_PD + i, ONE}}, {{S1_PD + i, ONE}});
    E.con(r + 19, V, {{S1_PD + i, ONE}}, {{S1_PD + i, ONE}}, {{S2_PD + i, ONE}});
    E.con(r + 20, V, {{S2_PD + i, ONE}, {AX + i, MINUS_TWO}}, {{K1, ONE}}, {{DX + i, ONE}});
    E.con(r + 21, V, {{S1_PD + i, ONE}}, {{AX + i, ONE}, {DX + i, MINUS_ONE}}, {{T2_PD + i, ONE}});
    E.con(r + 22, V, {{T2_PD + i, ONE}, {AY + i, MINUS_ONE}}, {{K1, ONE}}, {{DY + i, ONE}});
    // B' = bit ? C : B ;  A' = D   :256-304
    E.con(r + 23, V, {{CX + i, ONE}}, {{BIT + i, ONE}}, {{Z1 + i, ONE}});
    E.con(r + 24, V, {{BX + i, ONE}}, {{K1, ONE}, {BIT + i, MINUS_ONE}}, {{Z2 + i, ONE}});
    E.con(r + 25, V, {{Z1 + i, ONE}, {Z2 + i, ONE}}, {{K1, ONE}}, {{BX + 1 + i, ONE}});
    E.con(r + 26, V, {{CY + i, ONE}}, {{BIT + i, ONE}}, {{Z3 + i, ONE}});
    E.con(r + 27, V, {{BY + i, ONE}}, {{K1, ONE}, {BIT + i, MINUS_ONE}}, {{Z4 + i, ONE}});
    E.con(r + 28, V, {{Z3 + i, ONE}, {Z4 + i, ONE}}, {{K1, ONE}}, {{BY + 1 + i, ONE}});
    E.con(r + 29, V, {{BZ + i, ONE}}, {{K1, ONE}, {BIT + i, MINUS_ONE}}, {{BZ + 1 + i, ONE}});
    E.con(r + 30, V, {{AX + 1 + i, ONE}, {DX + i, MINUS_ONE}}, {{K1, ONE}}, {});
    E.con(r + 31, V, {{AY + 1 + i, ONE}, {DY + i, MINUS_ONE}}, {{K1, ONE}}, {});
  }
  E.con(R + PM_ONC - 2, V, {{QX, ONE}, {BX + n, MINUS_ONE}}, {{K1, ONE}}, {});
  E.con(R + PM_ONC - 1, V, {{QY, ONE}, {BY + n, MINUS_ONE}}, {{K1, ONE}}, {});
}

// out[j * per + k] = tmpl[k] moved to block j: rows += onc * j; columns below the sentinel += onv * j, the two codes above it
// become num_vars (constant 1) and num_vars + 1 (public input 0)
__global__ void __launch_bounds__(256) k_pm_emit(const vpin_coo_entry *tmpl, size_t per, size_t m, uint64_t onc, uint64_t onv, uint64_t num_vars,
                                                 vpin_coo_entry *out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per * m) return;
  size_t j = idx / per, k = idx % per;
  const uint4 *q = reinterpret_cast<const uint4 *>(tmpl + k);
  uint4 rc = __ldg(q), v0 = __ldg(q + 1), v1 = __ldg(q + 2);
  uint64_t r = ((uint64_t)rc.x | ((uint64_t)rc.y << 32)) + onc * j, c = (uint64_t)rc.z | ((uint64_t)rc.w << 32);
  c = c >= kColSentinel ? num_vars + (c - kColSentinel) : c + onv * j;
  uint4 *o = reinterpret_cast<uint4 *>(out + idx);
  o[0] = make_uint4((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)c, (uint32_t)(c >> 32));
  o[1] = v0;
  o[2] = v1;
}

__device__ __forceinline__ void stw(fl_t *W, size_t i, const fl_t &x) {
  uint4 *q = reinterpret_cast<uint4 *>(W + i);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ fl_t ld_raw32(const uint8_t *p) {  // 32 little-endian bytes -> raw limbs (16-byte aligned input)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fl_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
// ---- witness expansion (point_mult.rs:328-602; pa :667-685, pd :687-704) --------------------------------------------------
// The reference walks the 128 double-and-add steps in affine coordinates: two field inversions per step, 256 in a row per
// multiplication. Every witness value of step i is a function of the affine points A_i = 2^i P, B_i = (k mod 2^i) P (with its
// infinity flag), the bit, and the two inverses c_i = 1 / (bx_i - ax_i), c_pd_i = 1 / (2 ay_i). So:
//   k_pm_chain   one thread per multiplication: both chains in Jacobian coordinates, no inversion (10 multiplications per
//                doubling, 16 per addition actually taken);
//   k_pm_invert  one thread per multiplication: Montgomery's trick over its 256 Z coordinates (one exponentiation);
//   k_pm_affine  one thread per (multiplication, step): affine A_i, B_i and the 256 denominators;
//   k_pm_invert  again, on the denominators;
//   k_pm_fill    one thread per (multiplication, step): the 27 values of the step, written in Montgomery form where the
//                prover reads them.
// ~5 K sequential field multiplications per multiplication instead of ~45 K; the values are the same field elements, hence the
// same bytes. Degenerate additions (B_i = +-A_i) cannot occur for a point of large prime order (B_i = (k mod 2^i) P with
// 0 < k mod 2^i < 2^i), and the reference itself would divide by zero there. Inversion maps 0 to 0 like dalek's.
struct jac_t { fl_t X, Y, Z; };
__device__ __forceinline__ fl_t ldw(const fl_t *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1];
  fl_t r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
// dbl-2007-bl for y^2 = x^3 + a x + b
__device__ __forceinline__ jac_t jac_dbl(const jac_t &p, const fl_t &a) {
  fl_t XX = fl_sqr(p.X), YY = fl_sqr(p.Y), YYYY = fl_sqr(YY), ZZ = fl_sqr(p.Z);
  fl_t S = fl_dbl(fl_sub(fl_sub(fl_sqr(fl_add(p.X, YY)), XX), YYYY));
  fl_t M = fl_add(fl_add(fl_dbl(XX), XX), fl_mul(a, fl_sqr(ZZ)));
  jac_t r;
  r.X = fl_sub(fl_sqr(M), fl_dbl(S));
  fl_t y8 = fl_dbl(fl_dbl(fl_dbl(YYYY)));
  r.Y = fl_sub(fl_mul(M, fl_sub(S, r.X)), y8);
  r.Z = fl_sub(fl_sub(fl_sqr(fl_add(p.Y, p.Z)), YY), ZZ);
  return r;
}
// add-2007-bl (distinct points, neither at infinity)
__device__ __forceinline__ jac_t jac_add(const jac_t &p, const jac_t &q) {
  fl_t Z1Z1 = fl_sqr(p.Z), Z2Z2 = fl_sqr(q.Z);
  fl_t U1 = fl_mul(p.X, Z2Z2), U2 = fl_mul(q.X, Z1Z1);
  fl_t S1 = fl_mul(fl_mul(p.Y, q.Z), Z2Z2), S2 = fl_mul(fl_mul(q.Y, p.Z), Z1Z1);
  fl_t H = fl_sub(U2, U1);
  fl_t I = fl_sqr(fl_dbl(H));
  fl_t J = fl_mul(H, I);
  fl_t rr = fl_dbl(fl_sub(S2, S1));
  fl_t V = fl_mul(U1, I);
  jac_t r;
  r.X = fl_sub(fl_sub(fl_sqr(rr), J), fl_dbl(V));
  r.Y = fl_sub(fl_mul(rr, fl_sub(V, r.X)), fl_dbl(fl_mul(S1, J)));
  r.Z = fl_mul(fl_sub(fl_sub(fl_sqr(fl_add(p.Z, q.Z)), Z1Z1), Z2Z2), H);
  return r;
}
// chains[(j * 128 + i) * 2 + 0] = A_i, [.. + 1] = B_i (Z = 0 for the point at infinity); zs[j * 256 + 2 i + {0, 1}] = their Z
__global__ void __launch_bounds__(32) k_pm_chain(size_t m, const uint64_t *weights, const uint8_t *px, const uint8_t *py, fl_t a_pd, jac_t *chains,
                                                 fl_t *zs) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint64_t wl = weights[2 * j], wh = weights[2 * j + 1];
  jac_t A, B;
  A.X = fl_mul(ld_raw32(px + 32 * j), fl_r2());  // Scalar::from_bytes_mod_order: x R^2 / R reduces any 256-bit x
  A.Y = fl_mul(ld_raw32(py + 32 * j), fl_r2());
  A.Z = fl_one();
  B.X = B.Y = B.Z = fl_zero();
  bool b_inf = true;
  jac_t *out = chains + j * PM_N * 2;
  fl_t *z = zs + j * PM_N * 2;
#pragma unroll 1
  for (int i = 0; i < PM_N; i++) {
    stw(&out[2 * i].X, 0, A.X); stw(&out[2 * i].Y, 0, A.Y); stw(&out[2 * i].Z, 0, A.Z);
    stw(&out[2 * i + 1].X, 0, B.X); stw(&out[2 * i + 1].Y, 0, B.Y); stw(&out[2 * i + 1].Z, 0, B.Z);
    stw(z, 2 * i, A.Z); stw(z, 2 * i + 1, B.Z);
    if ((i < 64 ? wl >> i : wh >> (i - 64)) & 1) {
      if (b_inf) { B = A; b_inf = false; }
      else B = jac_add(B, A);
    }
    A = jac_dbl(A, a_pd);
  }
}
// v[j * per .. (j + 1) * per) := element-wise inverses (zeros stay zero); tmp: same size, prefix products
__global__ void __launch_bounds__(32) k_pm_invert(size_t m, int per, fl_t *v, fl_t *tmp) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  fl_t *x = v + j * per, *pre = tmp + j * per;
  fl_t run = fl_one();
#pragma unroll 1
  for (int k = 0; k < per; k++) {
    stw(pre, k, run);
    fl_t e = ldw(x + k);
    if (!fl_is_zero(e)) run = fl_mul(run, e);
  }
  fl_t inv = fl_invert(run);
#pragma unroll 1
  for (int k = per - 1; k >= 0; k--) {
    fl_t e = ldw(x + k);
    if (fl_is_zero(e)) continue;
    stw(x, k, fl_mul(inv, ldw(pre + k)));
    inv = fl_mul(inv, e);
  }
}
// affine coordinates of A_i, B_i from the inverted Z's; aff[(j * 128 + i) * 4 + {0..3}] = ax, ay, bx, by;
// den[j * 256 + 2 i + {0, 1}] = bx - ax, 2 ay
__global__ void __launch_bounds__(128) k_pm_affine(size_t total, const jac_t *chains, const fl_t *zinv, fl_t *aff, fl_t *den) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  fl_t xy[4];
#pragma unroll
  for (int w = 0; w < 2; w++) {
    const jac_t *p = chains + 2 * t + w;
    fl_t zi = ldw(zinv + 2 * t + w);  // 0 for the point at infinity -> (0, 0)
    fl_t zi2 = fl_sqr(zi);
    xy[2 * w] = fl_mul(ldw(&p->X), zi2);
    xy[2 * w + 1] = fl_mul(ldw(&p->Y), fl_mul(zi2, zi));
  }
#pragma unroll
  for (int k = 0; k < 4; k++) stw(aff, 4 * t + k, xy[k]);
  stw(den, 2 * t, fl_sub(xy[2], xy[0]));
  stw(den, 2 * t + 1, fl_dbl(xy[1]));
}
// the 27 witness values of step i of multiplication j (+ the block's header / trailer values from steps 0 and 127)
__global__ void __launch_bounds__(128) k_pm_fill(size_t total, const uint64_t *weights, const jac_t *chains, const fl_t *aff, const fl_t *dinv,
                                                 fl_t a_pd, fl_t *W) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const size_t j = t / PM_N;
  const int i = (int)(t % PM_N);
  const size_t V = PM_ONV * j;
  const uint64_t wl = weights[2 * j], wh = weights[2 * j + 1];
  const fl_t one = fl_one(), zero = fl_zero();
  const bool bit_set = ((i < 64 ? wl >> i : wh >> (i - 64)) & 1) != 0;
  fl_t ax = ldw(aff + 4 * t), ay = ldw(aff + 4 * t + 1), bx = ldw(aff + 4 * t + 2), by = ldw(aff + 4 * t + 3);
  const bool b_inf = fl_is_zero(ldw(&chains[2 * t + 1].Z));
  fl_t bz = b_inf ? one : zero, nbz1 = b_inf ? zero : one;
  fl_t c = ldw(dinv + 2 * t), c_pd = ldw(dinv + 2 * t + 1);
  if (i == 0) {
    fl_t a = zero;
    a.v[0] = (uint32_t)wl; a.v[1] = (uint32_t)(wl >> 32); a.v[2] = (uint32_t)wh; a.v[3] = (uint32_t)(wh >> 32);
    stw(W, V + A_SC, fl_to_mont(a));
    stw(W, V + AX, ax); stw(W, V + AY, ay); stw(W, V + BZ, one);
    stw(W, V + PX, ax); stw(W, V + PY, ay);
  }
  fl_t s1 = fl_mul(fl_sub(by, ay), c);
  fl_t s2 = fl_sqr(s1);
  fl_t t1 = fl_mul(fl_sub(fl_sub(s2, ax), bx), nbz1);
  fl_t t2 = fl_mul(ax, bz);
  fl_t cx = fl_add(t1, t2);
  fl_t s3 = fl_mul(s1, fl_sub(ax, cx));
  fl_t t3 = fl_mul(fl_sub(s3, ay), nbz1);
  fl_t t4 = fl_mul(ay, bz);
  fl_t cy = fl_add(t3, t4);
  stw(W, V + C_PA + i, c); stw(W, V + S1_PA + i, s1); stw(W, V + S2_PA + i, s2); stw(W, V + S3_PA + i, s3);
  stw(W, V + T1_PA + i, t1); stw(W, V + T2_PA + i, t2); stw(W, V + T3_PA + i, t3); stw(W, V + T4_PA + i, t4);
  stw(W, V + CX + i, cx); stw(W, V + CY + i, cy);
  fl_t t1_pd = fl_sqr(ax);
  fl_t s1_pd = fl_mul(fl_add(fl_add(fl_dbl(t1_pd), t1_pd), a_pd), c_pd);
  fl_t s2_pd = fl_sqr(s1_pd);
  fl_t dx = fl_sub(s2_pd, fl_dbl(ax));
  fl_t t2_pd = fl_mul(s1_pd, fl_sub(ax, dx));
  fl_t dy = fl_sub(t2_pd, ay);
  stw(W, V + C_PD + i, c_pd); stw(W, V + T1_PD + i, t1_pd); stw(W, V + S1_PD + i, s1_pd); stw(W, V + S2_PD + i, s2_pd);
  stw(W, V + T2_PD + i, t2_pd); stw(W, V + DX + i, dx); stw(W, V + DY + i, dy);
  // z1 = cx * bit, z2 = bx * (1 - bit), ... : multiplications by 0 / 1
  fl_t z1 = bit_set ? cx : zero, z2 = bit_set ? zero : bx, z3 = bit_set ? cy : zero, z4 = bit_set ? zero : by;
  fl_t nbx = bit_set ? cx : bx, nby = bit_set ? cy : by, nbz = bit_set ? zero : bz;
  stw(W, V + BIT + i, bit_set ? one : zero);
  stw(W, V + Z1 + i, z1); stw(W, V + Z2 + i, z2); stw(W, V + Z3 + i, z3); stw(W, V + Z4 + i, z4);
  stw(W, V + AX + 1 + i, dx); stw(W, V + AY + 1 + i, dy);
  stw(W, V + BX + 1 + i, nbx); stw(W, V + BY + 1 + i, nby); stw(W, V + BZ + 1 + i, nbz);
  if (i == PM_N - 1) { stw(W, V + QX, nbx); stw(W, V + QY, nby); }
}
// vars_para holds only the scalar a of every block; vars_input everything else (point_mult.rs:517-571)
__global__ void __launch_bounds__(256) k_pm_split(const fl_t *W, size_t nv, size_t blocks_end, fl_t *para, fl_t *input) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nv) return;
  const uint4 *q = reinterpret_cast<const uint4 *>(W + k);
  uint4 a = q[0], b = q[1], z = make_uint4(0, 0, 0, 0);
  bool is_a = k < blocks_end && (k % PM_ONV) == (size_t)A_SC;
  uint4 *p = reinterpret_cast<uint4 *>(para + k), *in = reinterpret_cast<uint4 *>(input + k);
  p[0] = is_a ? a : z; p[1] = is_a ? b : z;
  in[0] = is_a ? z : a; in[1] = is_a ? z : b;
}
}  // namespace

std::unique_ptr<Instance> build_point_mult_dev(Ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32, const uint8_t *py32,
                                               uint64_t dims[4], fl_t *d_para, fl_t *d_input, fl_t *d_vars, size_t padded, uint8_t *inputs32) {
  point_mult_dims(m, dims);
  cudaStream_t st = ctx->st;
  const size_t nv = dims[1];
  VPIN_REQUIRE(m > 0 && padded >= nv, VPIN_ERR_SIZE_MISMATCH, "point_mult: assignment buffers too small");
  static const uint8_t a_pd_byte[32] = {157, 27, 50, 101, 63, 42, 38, 142, 68, 159, 245, 15, 16, 47, 75, 58,
                                        203, 87, 15, 3, 219, 183, 77, 94, 64, 118, 147, 233, 124, 16, 184, 7};  // :341
  fl_t a_pd = from_bytes_mod_order(a_pd_byte);
  fl_to_bytes(a_pd, inputs32);
  // ---- witness ----
  DevVec<uint64_t> d_w(2 * m, st);
  DevVec<uint8_t> d_px(32 * m, st), d_py(32 * m, st);
  d_w.upload(weights_lo_hi, 2 * m);
  d_px.upload(px32, 32 * m);
  d_py.upload(py32, 32 * m);
  VPIN_CUDA(cudaMemsetAsync(d_vars, 0, padded * sizeof(fl_t), st));
  VPIN_CUDA(cudaMemsetAsync(d_para, 0, padded * sizeof(fl_t), st));
  VPIN_CUDA(cudaMemsetAsync(d_input, 0, padded * sizeof(fl_t), st));
  {
    const size_t steps = m * PM_N;
    DevVec<jac_t> chains(2 * steps, st);
    DevVec<fl_t> zs(2 * steps, st), tmp(2 * steps, st), aff(4 * steps, st), den(2 * steps, st);
    const unsigned jb = (unsigned)((m + 31) / 32), sb = (unsigned)((steps + 127) / 128);
    ++g_kernel_launches, k_pm_chain<<<jb, 32, 0, st>>>(m, d_w.p, d_px.p, d_py.p, a_pd, chains.p, zs.p);
    ++g_kernel_launches, k_pm_invert<<<jb, 32, 0, st>>>(m, 2 * (int)PM_N, zs.p, tmp.p);
    ++g_kernel_launches, k_pm_affine<<<sb, 128, 0, st>>>(steps, chains.p, zs.p, aff.p, den.p);
    ++g_kernel_launches, k_pm_invert<<<jb, 32, 0, st>>>(m, 2 * (int)PM_N, den.p, tmp.p);
    ++g_kernel_launches, k_pm_fill<<<sb, 128, 0, st>>>(steps, d_w.p, chains.p, aff.p, den.p, a_pd, d_vars);
  }
  ++g_kernel_launches, k_pm_split<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(d_vars, nv, PM_ONV * m, d_para, d_input);
  // ---- constraint system ----
  Emitter E;
  E.num_vars = kColSentinel;
  emit_point_mult_block(E, 0, 0);
  std::unique_ptr<Instance> inst;
  {
    DevVec<vpin_coo_entry> tmpl[3], full[3];
    for (int k = 0; k < 3; k++) {
      size_t per = E.M[k].size();
      tmpl[k].alloc(per, st);
      tmpl[k].upload(E.M[k].data(), per);
      full[k].alloc(per * m, st);
      ++g_kernel_launches, k_pm_emit<<<(unsigned)((per * m + 255) / 256), 256, 0, st>>>(tmpl[k].p, per, m, PM_ONC, PM_ONV, nv, full[k].p);
    }
    ctx->sync();  // the template vectors of E go out of use here
    inst = instance_create(ctx, dims[0], dims[1], dims[2], full[0].p, E.M[0].size() * m, full[1].p, E.M[1].size() * m, full[2].p,
                           E.M[2].size() * m, true);
  }
  return inst;
}

// host-buffer variant (the JSON-driven flow of proof_point_mult.rs): the same device build, assignments downloaded as
// canonical bytes
std::unique_ptr<Instance> build_point_mult(Ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32, const uint8_t *py32,
                                           uint64_t dims[4], uint8_t *vars_para32, uint8_t *vars_input32, uint8_t *vars32,
                                           uint8_t *inputs32) {
  point_mult_dims(m, dims);
  size_t nv = dims[1];
  cudaStream_t st = ctx->st;
  DevVec<fl_t> d_para(nv, st), d_input(nv, st), d_vars(nv, st);
  auto inst = build_point_mult_dev(ctx, m, weights_lo_hi, px32, py32, dims, d_para.p, d_input.p, d_vars.p, nv, inputs32);
  uint8_t *outs[3] = {vars_para32, vars_input32, vars32};
  fl_t *srcs[3] = {d_para.p, d_input.p, d_vars.p};
  for (int k = 0; k < 3; k++) {
    launch_from_mont(srcs[k], nv, srcs[k], st);
    VPIN_CUDA(cudaMemcpyAsync(outs[k], srcs[k], nv * sizeof(fl_t), cudaMemcpyDeviceToHost, st));
  }
  ctx->sync();
  return inst;
}

}  // namespace vpin
