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

std::unique_ptr<Instance> build_point_mult(Ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32, const uint8_t *py32,
                                           uint64_t dims[4], uint8_t *vars_para32, uint8_t *vars_input32, uint8_t *vars32,
                                           uint8_t *inputs32) {
  point_mult_dims(m, dims);
  const long n = 128;
  const size_t onc = 27 * n + 8, onv = 27 * n + 10;
  Emitter E;
  E.num_vars = dims[1];
  // variable layout of one block (point_mult.rs:517-601)
  const long BIT = 0, A_SC = n, AX = n + 1, AY = 2 * n + 2, BX = 3 * n + 3, BY = 4 * n + 4, BZ = 5 * n + 5, CX = 6 * n + 6, CY = 7 * n + 6,
             DX = 8 * n + 6, DY = 9 * n + 6, QX = 10 * n + 6, QY = 10 * n + 7, PX = 10 * n + 8, PY = 10 * n + 9, C_PA = 10 * n + 10,
             S1_PA = 11 * n + 10, S2_PA = 12 * n + 10, S3_PA = 13 * n + 10, T1_PA = 14 * n + 10, T2_PA = 15 * n + 10, T3_PA = 16 * n + 10,
             T4_PA = 17 * n + 10, C_PD = 18 * n + 10, T1_PD = 19 * n + 10, S1_PD = 20 * n + 10, S2_PD = 21 * n + 10, T2_PD = 22 * n + 10,
             Z1 = 23 * n + 10, Z2 = 24 * n + 10, Z3 = 25 * n + 10, Z4 = 26 * n + 10;
  uint8_t pow2[128][32];
  {
    fl_t tb = fl_one(), two = fl_one() + fl_one();
    for (int i = 0; i < n; i++) { fl_to_bytes(tb, pow2[i]); tb = tb * two; }
  }
  for (size_t j = 0; j < m; j++) {  // :85-322
    size_t R = onc * j, V = onv * j;
    // sum_i 2^i bit_i * 1 = a
    for (long i = 0; i < n; i++) {
      vpin_coo_entry e;
      e.row = R; e.col = V + i;
      memcpy(e.val, pow2[i], 32);
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
    E.con(R + onc - 2, V, {{QX, ONE}, {BX + n, MINUS_ONE}}, {{K1, ONE}}, {});
    E.con(R + onc - 1, V, {{QY, ONE}, {BY + n, MINUS_ONE}}, {{K1, ONE}}, {});
  }

  // ---- witness expansion (:328-602, pa :667-685, pd :687-704) ----
  static const uint8_t a_pd_byte[32] = {157, 27, 50, 101, 63, 42, 38, 142, 68, 159, 245, 15, 16, 47, 75, 58,
                                        203, 87, 15, 3, 219, 183, 77, 94, 64, 118, 147, 233, 124, 16, 184, 7};  // :341
  fl_t a_pd = from_bytes_mod_order(a_pd_byte);
  size_t nv = dims[1];
  std::vector<fl_t> W(nv, fl_zero());  // the full assignment `vars`
  fl_t one = fl_one(), zero = fl_zero(), two = one + one, three = two + one;
  std::vector<fl_t> ax(m), ay(m), bx(m, zero), by(m, zero), bz(m, one), inv(2 * m);
  for (size_t j = 0; j < m; j++) {
    size_t V = onv * j;
    fl_t a = fl_zero();
    a.v[0] = (uint32_t)weights_lo_hi[2 * j]; a.v[1] = (uint32_t)(weights_lo_hi[2 * j] >> 32);
    a.v[2] = (uint32_t)weights_lo_hi[2 * j + 1]; a.v[3] = (uint32_t)(weights_lo_hi[2 * j + 1] >> 32);
    W[V + A_SC] = fl_to_mont(a);
    ax[j] = from_bytes_mod_order(px32 + 32 * j);
    ay[j] = from_bytes_mod_order(py32 + 32 * j);
    W[V + AX] = ax[j]; W[V + AY] = ay[j]; W[V + BX] = zero; W[V + BY] = zero; W[V + BZ] = one;
    W[V + PX] = ax[j]; W[V + PY] = ay[j];
  }
  for (long i = 0; i < n; i++) {
    for (size_t j = 0; j < m; j++) { inv[2 * j] = bx[j] - ax[j]; inv[2 * j + 1] = two * ay[j]; }
    batch_invert(inv);
    for (size_t j = 0; j < m; j++) {
      size_t V = onv * j;
      uint64_t wl = weights_lo_hi[2 * j], wh = weights_lo_hi[2 * j + 1];
      fl_t bit = ((i < 64 ? wl >> i : wh >> (i - 64)) & 1) ? one : zero;
      fl_t c = inv[2 * j];
      fl_t s1 = (by[j] - ay[j]) * c;
      fl_t s2 = s1 * s1;
      fl_t t1 = (s2 - ax[j] - bx[j]) * (one - bz[j]);
      fl_t t2 = ax[j] * bz[j];
      fl_t cx = t1 + t2;
      fl_t s3 = s1 * (ax[j] - cx);
      fl_t t3 = (s3 - ay[j]) * (one - bz[j]);
      fl_t t4 = ay[j] * bz[j];
      fl_t cy = t3 + t4;
      fl_t c_pd = inv[2 * j + 1];
      fl_t t1_pd = ax[j] * ax[j];
      fl_t s1_pd = (three * t1_pd + a_pd) * c_pd;
      fl_t s2_pd = s1_pd * s1_pd;
      fl_t dx = s2_pd - two * ax[j];
      fl_t t2_pd = s1_pd * (ax[j] - dx);
      fl_t dy = t2_pd - ay[j];
      fl_t z1 = cx * bit, z2 = bx[j] * (one - bit), z3 = cy * bit, z4 = by[j] * (one - bit);
      fl_t nbx = z1 + z2, nby = z3 + z4, nbz = bz[j] * (one - bit);
      W[V + BIT + i] = bit;
      W[V + AX + 1 + i] = dx; W[V + AY + 1 + i] = dy;
      W[V + BX + 1 + i] = nbx; W[V + BY + 1 + i] = nby; W[V + BZ + 1 + i] = nbz;
      W[V + CX + i] = cx; W[V + CY + i] = cy; W[V + DX + i] = dx; W[V + DY + i] = dy;
      W[V + C_PA + i] = c; W[V + S1_PA + i] = s1; W[V + S2_PA + i] = s2; W[V + S3_PA + i] = s3;
      W[V + T1_PA + i] = t1; W[V + T2_PA + i] = t2; W[V + T3_PA + i] = t3; W[V + T4_PA + i] = t4;
      W[V + C_PD + i] = c_pd; W[V + T1_PD + i] = t1_pd; W[V + S1_PD + i] = s1_pd; W[V + S2_PD + i] = s2_pd; W[V + T2_PD + i] = t2_pd;
      W[V + Z1 + i] = z1; W[V + Z2 + i] = z2; W[V + Z3 + i] = z3; W[V + Z4 + i] = z4;
      ax[j] = dx; ay[j] = dy; bx[j] = nbx; by[j] = nby; bz[j] = nbz;
    }
  }
  for (size_t j = 0; j < m; j++) { W[onv * j + QX] = bx[j]; W[onv * j + QY] = by[j]; }
  // vars_para holds only the scalar a; vars_input everything else (:517-571)
  memset(vars_para32, 0, 32 * nv);
  for (size_t k = 0; k < nv; k++) fl_to_bytes(W[k], vars32 + 32 * k);
  memcpy(vars_input32, vars32, 32 * nv);
  for (size_t j = 0; j < m; j++) {
    size_t k = onv * j + A_SC;
    memcpy(vars_para32 + 32 * k, vars32 + 32 * k, 32);
    memset(vars_input32 + 32 * k, 0, 32);
  }
  fl_to_bytes(a_pd, inputs32);
  return instance_create(ctx, dims[0], dims[1], dims[2], E.M[0].data(), E.M[0].size(), E.M[1].data(), E.M[1].size(), E.M[2].data(),
                         E.M[2].size());
}

}  // namespace vpin
