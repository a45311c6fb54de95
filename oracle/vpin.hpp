// TEST INFRASTRUCTURE ONLY (CPU oracle). Not linked into the product library.
//
// Restates vPIN's R1CS builders and the driver flow around my_lib_prove:
//   VP/point_addition.rs:5-326  (10 constraints / 15 variables per EC point addition)
//   VP/point_mult.rs:7-704      (27n+8 constraints / 27n+10 variables per 128-bit double-and-add)
//   VP/proof_point_add.rs:23-113 == VP/proof_point_mult.rs (gens -> encode -> 3 commits -> prove -> verify)
// Witness inputs are what VP/load_data.rs / load_data_add.rs parse from the Python side's JSON files:
// 32-byte little-endian coordinates and u128 weights.
#pragma once
#include "spartan.hpp"

namespace orc {

struct BuiltInstance {
  size_t num_cons, num_vars, num_inputs, num_non_zero_entries;
  std::vector<CooEntry> A, B, C;
  std::vector<std::array<uint8_t, 32>> vars_para, vars_input, vars, inputs;  // canonical LE bytes, unpadded
};

static inline Fl fl_from_bytes_mod_order(const uint8_t b[32]) {  // dalek Scalar::from_bytes_mod_order
  uint8_t w[64];
  memcpy(w, b, 32);
  memset(w + 32, 0, 32);
  return fl_from_bytes_wide(w);
}
static inline std::array<uint8_t, 32> fl_bytes(const Fl &x) { std::array<uint8_t, 32> o; fl_to_bytes(x, o.data()); return o; }
static inline void coo_push(std::vector<CooEntry> &M, size_t row, size_t col, const std::array<uint8_t, 32> &v) {
  CooEntry e;
  e.row = row; e.col = col;
  memcpy(e.val, v.data(), 32);
  M.push_back(e);
}

// VP/point_addition.rs:5-326
static inline BuiltInstance build_point_add(size_t n, const uint8_t *px_b, const uint8_t *py_b, const uint8_t *rx_b,
                                            const uint8_t *ry_b, const int64_t *rz_flag) {
  BuiltInstance bi;
  size_t p1, p2, p3;  // :38-65
  if (n < 780) { p1 = 2; p2 = 25; p3 = 3; }
  else if (n >= 500 && n < 780) { p1 = 2; p2 = 25; p3 = 3; }
  else if (n > 2130 && n < 2150) { p1 = 5; p2 = 30; p3 = 5; }
  else if (n > 2149 && n < 2450) { p1 = 3; p2 = 30; p3 = 5; }
  else if (n > 5000 && n < 8000) { p1 = 3; p2 = 20; p3 = 5; }
  else { p1 = 5; p2 = 30; p3 = 5; }
  bi.num_cons = 10 * n;
  bi.num_vars = 15 * n + 1;
  bi.num_inputs = 0;
  bi.num_non_zero_entries = p1 * (p2 / p3) * n;
  size_t nv = bi.num_vars;
  auto one = fl_bytes(fl_one()), minus_one = fl_bytes(fl_neg(fl_one()));
  auto &A = bi.A, &B = bi.B, &C = bi.C;
  for (size_t i = 0; i < n; i++) {  // :81-151
    size_t r = 10 * i, v = 15 * i;
    coo_push(A, r + 0, v + 0, one); coo_push(B, r + 0, v + 1, one); coo_push(B, r + 0, v + 2, minus_one); coo_push(C, r + 0, nv, one);
    coo_push(A, r + 1, v + 3, one); coo_push(A, r + 1, v + 4, minus_one); coo_push(B, r + 1, v + 0, one); coo_push(C, r + 1, v + 6, one);
    coo_push(A, r + 2, v + 6, one); coo_push(B, r + 2, v + 6, one); coo_push(C, r + 2, v + 7, one);
    coo_push(A, r + 3, v + 7, one); coo_push(A, r + 3, v + 2, minus_one); coo_push(A, r + 3, v + 1, minus_one);
    coo_push(B, r + 3, nv, one); coo_push(B, r + 3, v + 5, minus_one); coo_push(C, r + 3, v + 9, one);
    coo_push(A, r + 4, v + 2, one); coo_push(B, r + 4, v + 5, one); coo_push(C, r + 4, v + 10, one);
    coo_push(A, r + 5, v + 9, one); coo_push(A, r + 5, v + 10, one); coo_push(B, r + 5, nv, one); coo_push(C, r + 5, v + 13, one);
    coo_push(A, r + 6, v + 6, one); coo_push(B, r + 6, v + 2, one); coo_push(B, r + 6, v + 13, minus_one); coo_push(C, r + 6, v + 8, one);
    coo_push(A, r + 7, v + 8, one); coo_push(A, r + 7, v + 4, minus_one); coo_push(B, r + 7, nv, one);
    coo_push(B, r + 7, v + 5, minus_one); coo_push(C, r + 7, v + 11, one);
    coo_push(A, r + 8, v + 4, one); coo_push(B, r + 8, v + 5, one); coo_push(C, r + 8, v + 12, one);
    coo_push(A, r + 9, v + 11, one); coo_push(A, r + 9, v + 12, one); coo_push(B, r + 9, nv, one); coo_push(C, r + 9, v + 14, one);
  }
  std::array<uint8_t, 32> zero_b = fl_bytes(fl_zero());
  bi.vars_para.assign(nv, zero_b);  // :223 all zero
  bi.vars_input.assign(nv, zero_b);
  Fl one_s = fl_one();
  for (size_t i = 0; i < n; i++) {  // :164-267
    Fl px = fl_from_bytes_mod_order(px_b + 32 * i), py = fl_from_bytes_mod_order(py_b + 32 * i);
    Fl rx = fl_from_bytes_mod_order(rx_b + 32 * i), ry = fl_from_bytes_mod_order(ry_b + 32 * i);
    Fl rz = rz_flag[i] == 0 ? fl_zero() : fl_one();
    Fl c = fl_invert(rx - px);
    Fl s1 = (ry - py) * c;
    Fl s2 = s1 * s1;
    Fl t1 = (s2 - px - rx) * (one_s - rz);
    Fl t2 = px * rz;
    Fl x3 = (t1 + t2) * one_s;
    Fl s3 = s1 * (px - x3);
    Fl t3 = (s3 - py) * (one_s - rz);
    Fl t4 = py * rz;
    Fl y3 = (t3 + t4) * one_s;
    Fl vals[15] = {c, rx, px, ry, py, rz, s1, s2, s3, t1, t2, t3, t4, x3, y3};
    for (int k = 0; k < 15; k++) bi.vars_input[15 * i + k] = fl_bytes(vals[k]);
  }
  bi.vars = bi.vars_input;
  return bi;
}

// VP/point_mult.rs:7-704. weights are u128 as (lo, hi) pairs (VP/load_data.rs:20-23); n = 128 (:62)
static inline BuiltInstance build_point_mult(size_t m, const uint64_t *weights_lo_hi, const uint8_t *px_b,
                                             const uint8_t *py_b, size_t n = 128) {
  BuiltInstance bi;
  size_t p1, p2, p3;  // :27-56
  if (m == 50) { p1 = 100; p2 = 2; p3 = 80; }
  else if (m == 210) { p1 = 300; p2 = 2; p3 = 20; }
  else if (m == 240) { p1 = 300; p2 = 4; p3 = 20; }
  else if (m < 660) { p1 = 100; p2 = 2; p3 = 40; }
  else if (m == 6000) { p1 = 250; p2 = 2; p3 = 20; }
  else { p1 = 350; p2 = 2; p3 = 20; }
  size_t onc = 27 * n + 8, onv = n + 10 + n * 26;
  bi.num_cons = onc * m;
  bi.num_vars = onv * m + 1;
  bi.num_inputs = 1;
  bi.num_non_zero_entries = p1 * (p2 * n + p3 * m);
  size_t nv = bi.num_vars;
  Fl one_s = fl_one();
  auto one = fl_bytes(one_s), two = fl_bytes(one_s + one_s), three = fl_bytes(one_s + one_s + one_s);
  auto minus_one = fl_bytes(fl_neg(one_s)), minus_two = fl_bytes(fl_zero() - one_s - one_s);
  auto &A = bi.A, &B = bi.B, &C = bi.C;
  for (size_t j = 0; j < m; j++) {  // :85-322
    size_t R = onc * j, V = onv * j;
    Fl two_base = one_s;
    for (size_t i = 0; i < n; i++) { coo_push(A, R, i + V, fl_bytes(two_base)); two_base = two_base * fl_from_u64(2); }
    coo_push(B, R, nv, one);
    coo_push(C, R, n + V, one);
    for (size_t i = 1; i < n + 1; i++) { coo_push(A, i + R, i - 1 + V, one); coo_push(B, i + R, i - 1 + V, one); coo_push(C, i + R, i - 1 + V, one); }
    coo_push(A, n + 1 + R, n + 1 + V, one); coo_push(A, n + 1 + R, 10 * n + 8 + V, minus_one); coo_push(B, n + 1 + R, nv, one);
    coo_push(A, n + 2 + R, 2 * n + 2 + V, one); coo_push(A, n + 2 + R, 10 * n + 9 + V, minus_one); coo_push(B, n + 2 + R, nv, one);
    coo_push(A, n + 3 + R, 3 * n + 3 + V, one); coo_push(B, n + 3 + R, nv, one);
    coo_push(A, n + 4 + R, 4 * n + 4 + V, one); coo_push(B, n + 4 + R, nv, one);
    coo_push(A, n + 5 + R, 5 * n + 5 + V, one); coo_push(A, n + 5 + R, nv, minus_one); coo_push(B, n + 5 + R, nv, one);
    for (size_t i = 0; i < n; i++) {
      size_t r = n + i * 26 + R;
      // PA :129-198
      coo_push(A, r + 6, 10 * n + 10 + i + V, one); coo_push(B, r + 6, 3 * n + 3 + i + V, one);
      coo_push(B, r + 6, n + 1 + i + V, minus_one); coo_push(C, r + 6, nv, one);
      coo_push(A, r + 7, 4 * n + 4 + i + V, one); coo_push(A, r + 7, 2 * n + 2 + i + V, minus_one);
      coo_push(B, r + 7, 10 * n + 10 + i + V, one); coo_push(C, r + 7, 11 * n + 10 + i + V, one);
      coo_push(A, r + 8, 11 * n + 10 + i + V, one); coo_push(B, r + 8, 11 * n + 10 + i + V, one); coo_push(C, r + 8, 12 * n + 10 + i + V, one);
      coo_push(A, r + 9, 12 * n + 10 + i + V, one); coo_push(A, r + 9, n + 1 + i + V, minus_one); coo_push(A, r + 9, 3 * n + 3 + i + V, minus_one);
      coo_push(B, r + 9, nv, one); coo_push(B, r + 9, 5 * n + 5 + i + V, minus_one); coo_push(C, r + 9, 14 * n + 10 + i + V, one);
      coo_push(A, r + 10, n + 1 + i + V, one); coo_push(B, r + 10, 5 * n + 5 + i + V, one); coo_push(C, r + 10, 15 * n + 10 + i + V, one);
      coo_push(A, r + 11, 14 * n + 10 + i + V, one); coo_push(A, r + 11, 15 * n + 10 + i + V, one);
      coo_push(B, r + 11, nv, one); coo_push(C, r + 11, 6 * n + 6 + i + V, one);
      coo_push(A, r + 12, 11 * n + 10 + i + V, one); coo_push(B, r + 12, n + 1 + i + V, one);
      coo_push(B, r + 12, 6 * n + 6 + i + V, minus_one); coo_push(C, r + 12, 13 * n + 10 + i + V, one);
      coo_push(A, r + 13, 13 * n + 10 + i + V, one); coo_push(A, r + 13, 2 * n + 2 + i + V, minus_one);
      coo_push(B, r + 13, nv, one); coo_push(B, r + 13, 5 * n + 5 + i + V, minus_one); coo_push(C, r + 13, 16 * n + 10 + i + V, one);
      coo_push(A, r + 14, 2 * n + 2 + i + V, one); coo_push(B, r + 14, 5 * n + 5 + i + V, one); coo_push(C, r + 14, 17 * n + 10 + i + V, one);
      coo_push(A, r + 15, 16 * n + 10 + i + V, one); coo_push(A, r + 15, 17 * n + 10 + i + V, one);
      coo_push(B, r + 15, nv, one); coo_push(C, r + 15, 7 * n + 6 + i + V, one);
      // PD :206-250
      coo_push(A, r + 16, 18 * n + 10 + i + V, one); coo_push(B, r + 16, 2 * n + 2 + i + V, two); coo_push(C, r + 16, nv, one);
      coo_push(A, r + 17, n + 1 + i + V, one); coo_push(B, r + 17, n + 1 + i + V, one); coo_push(C, r + 17, 19 * n + 10 + i + V, one);
      coo_push(A, r + 18, 19 * n + 10 + i + V, three); coo_push(A, r + 18, nv + 1, one);
      coo_push(B, r + 18, 18 * n + 10 + i + V, one); coo_push(C, r + 18, 20 * n + 10 + i + V, one);
      coo_push(A, r + 19, 20 * n + 10 + i + V, one); coo_push(B, r + 19, 20 * n + 10 + i + V, one); coo_push(C, r + 19, 21 * n + 10 + i + V, one);
      coo_push(A, r + 20, 21 * n + 10 + i + V, one); coo_push(A, r + 20, n + 1 + i + V, minus_two);
      coo_push(B, r + 20, nv, one); coo_push(C, r + 20, 8 * n + 6 + i + V, one);
      coo_push(A, r + 21, 20 * n + 10 + i + V, one); coo_push(B, r + 21, n + 1 + i + V, one);
      coo_push(B, r + 21, 8 * n + 6 + i + V, minus_one); coo_push(C, r + 21, 22 * n + 10 + i + V, one);
      coo_push(A, r + 22, 22 * n + 10 + i + V, one); coo_push(A, r + 22, 2 * n + 2 + i + V, minus_one);
      coo_push(B, r + 22, nv, one); coo_push(C, r + 22, 9 * n + 6 + i + V, one);
      // select :256-304
      coo_push(A, r + 23, 6 * n + 6 + i + V, one); coo_push(B, r + 23, i + V, one); coo_push(C, r + 23, 23 * n + 10 + i + V, one);
      coo_push(A, r + 24, 3 * n + 3 + i + V, one); coo_push(B, r + 24, nv, one); coo_push(B, r + 24, i + V, minus_one);
      coo_push(C, r + 24, 24 * n + 10 + i + V, one);
      coo_push(A, r + 25, 23 * n + 10 + i + V, one); coo_push(A, r + 25, 24 * n + 10 + i + V, one);
      coo_push(B, r + 25, nv, one); coo_push(C, r + 25, 3 * n + 4 + i + V, one);
      coo_push(A, r + 26, 7 * n + 6 + i + V, one); coo_push(B, r + 26, i + V, one); coo_push(C, r + 26, 25 * n + 10 + i + V, one);
      coo_push(A, r + 27, 4 * n + 4 + i + V, one); coo_push(B, r + 27, nv, one); coo_push(B, r + 27, i + V, minus_one);
      coo_push(C, r + 27, 26 * n + 10 + i + V, one);
      coo_push(A, r + 28, 25 * n + 10 + i + V, one); coo_push(A, r + 28, 26 * n + 10 + i + V, one);
      coo_push(B, r + 28, nv, one); coo_push(C, r + 28, 4 * n + 5 + i + V, one);
      coo_push(A, r + 29, 5 * n + 5 + i + V, one); coo_push(B, r + 29, nv, one); coo_push(B, r + 29, i + V, minus_one);
      coo_push(C, r + 29, 5 * n + 6 + i + V, one);
      coo_push(A, r + 30, n + 2 + i + V, one); coo_push(A, r + 30, 8 * n + 6 + i + V, minus_one); coo_push(B, r + 30, nv, one);
      coo_push(A, r + 31, 2 * n + 3 + i + V, one); coo_push(A, r + 31, 9 * n + 6 + i + V, minus_one); coo_push(B, r + 31, nv, one);
    }
    coo_push(A, onc - 2 + R, 10 * n + 6 + V, one); coo_push(A, onc - 2 + R, 3 * n + 3 + n + V, minus_one); coo_push(B, onc - 2 + R, nv, one);
    coo_push(A, onc - 1 + R, 10 * n + 7 + V, one); coo_push(A, onc - 1 + R, 4 * n + 4 + n + V, minus_one); coo_push(B, onc - 1 + R, nv, one);
  }

  // witness :328-602
  static const uint8_t a_pd_byte[32] = {157, 27, 50, 101, 63, 42, 38, 142, 68, 159, 245, 15, 16, 47, 75, 58,
                                        203, 87, 15, 3, 219, 183, 77, 94, 64, 118, 147, 233, 124, 16, 184, 7};
  Fl a_pd = fl_from_bytes_mod_order(a_pd_byte);
  Fl two_s = one_s + one_s, three_s = two_s + one_s;
  std::array<uint8_t, 32> zero_b = fl_bytes(fl_zero());
  bi.vars_para.assign(nv, zero_b);
  bi.vars_input.assign(nv, zero_b);
  bi.vars.assign(nv, zero_b);
  for (size_t j = 0; j < m; j++) {
    size_t V = onv * j;
    uint64_t wlo = weights_lo_hi[2 * j], whi = weights_lo_hi[2 * j + 1];
    uint64_t raw[4] = {wlo, whi, 0, 0};
    Fl a = fl_from_raw(raw);
    Fl px = fl_from_bytes_mod_order(px_b + 32 * j), py = fl_from_bytes_mod_order(py_b + 32 * j);
    Fl ax_prev = px, ay_prev = py, bx_prev = fl_zero(), by_prev = fl_zero(), bz_prev = one_s;
    auto setv = [&](std::vector<std::array<uint8_t, 32>> &dst, size_t idx, const Fl &x) { dst[idx + V] = fl_bytes(x); };
    auto set_in = [&](size_t idx, const Fl &x) { setv(bi.vars_input, idx, x); setv(bi.vars, idx, x); };
    bi.vars_para[n + V] = fl_bytes(a);
    bi.vars[n + V] = fl_bytes(a);
    set_in(n + 1, px); set_in(2 * n + 2, py); set_in(3 * n + 3, bx_prev); set_in(4 * n + 4, by_prev); set_in(5 * n + 5, bz_prev);
    for (size_t i = 0; i < n; i++) {
      bool bitv = i < 64 ? (wlo >> i) & 1 : (whi >> (i - 64)) & 1;
      Fl bit = bitv ? one_s : fl_zero();
      // pa :667-685
      Fl c = fl_invert(bx_prev - ax_prev);
      Fl s1 = (by_prev - ay_prev) * c;
      Fl s2 = s1 * s1;
      Fl t1 = (s2 - ax_prev - bx_prev) * (one_s - bz_prev);
      Fl t2 = ax_prev * bz_prev;
      Fl cx = (t1 + t2) * one_s;
      Fl s3 = s1 * (ax_prev - cx);
      Fl t3 = (s3 - ay_prev) * (one_s - bz_prev);
      Fl t4 = ay_prev * bz_prev;
      Fl cy = (t3 + t4) * one_s;
      // pd :687-704
      Fl c_pd = fl_invert(two_s * ay_prev);
      Fl t1_pd = ax_prev * ax_prev;
      Fl s1_pd = (three_s * t1_pd + a_pd) * c_pd;
      Fl s2_pd = s1_pd * s1_pd;
      Fl dx = s2_pd - two_s * ax_prev;
      Fl t2_pd = s1_pd * (ax_prev - dx);
      Fl dy = t2_pd - ay_prev;
      Fl z1 = cx * bit, z2 = bx_prev * (one_s - bit);
      Fl bx = z1 + z2;
      Fl z3 = cy * bit, z4 = by_prev * (one_s - bit);
      Fl by = z3 + z4;
      Fl bz = bz_prev * (one_s - bit);
      set_in(i, bit);
      set_in(n + 2 + i, dx); set_in(2 * n + 3 + i, dy);
      set_in(3 * n + 4 + i, bx); set_in(4 * n + 5 + i, by); set_in(5 * n + 6 + i, bz);
      set_in(6 * n + 6 + i, cx); set_in(7 * n + 6 + i, cy); set_in(8 * n + 6 + i, dx); set_in(9 * n + 6 + i, dy);
      set_in(10 * n + 10 + i, c); set_in(11 * n + 10 + i, s1); set_in(12 * n + 10 + i, s2); set_in(13 * n + 10 + i, s3);
      set_in(14 * n + 10 + i, t1); set_in(15 * n + 10 + i, t2); set_in(16 * n + 10 + i, t3); set_in(17 * n + 10 + i, t4);
      set_in(18 * n + 10 + i, c_pd); set_in(19 * n + 10 + i, t1_pd); set_in(20 * n + 10 + i, s1_pd);
      set_in(21 * n + 10 + i, s2_pd); set_in(22 * n + 10 + i, t2_pd);
      set_in(23 * n + 10 + i, z1); set_in(24 * n + 10 + i, z2); set_in(25 * n + 10 + i, z3); set_in(26 * n + 10 + i, z4);
      ax_prev = dx; ay_prev = dy; bx_prev = bx; by_prev = by; bz_prev = bz;
    }
    set_in(10 * n + 6, bx_prev); set_in(10 * n + 7, by_prev); set_in(10 * n + 8, px); set_in(10 * n + 9, py);
  }
  bi.inputs.push_back(fl_bytes(a_pd));
  return bi;
}

// SP/lib.rs:78-120 (Assignment::new + pad)
static inline bool assignment_new(const std::vector<std::array<uint8_t, 32>> &b, size_t pad_to, FlVec *out) {
  out->clear();
  for (auto &x : b) {
    Fl v;
    if (!fl_from_bytes(x.data(), &v)) return false;
    out->push_back(v);
  }
  if (pad_to > out->size()) out->resize(pad_to, fl_zero());
  return true;
}

struct FlowResult {
  std::vector<uint8_t> proof;        // bincode(SNARK)
  std::vector<uint8_t> comm;         // bincode(R1CSCommitment) = ComputationCommitment
  PolyCommitment comm_vars_para, comm_vars_input, comm_vars;
  bool verified;
  PhaseTimes times;
  double ms_gens, ms_commits, ms_verify;
};

// VP/proof_point_add.rs:39-107. seed_q / seed_p = init_randomness of RandomTape::new(&[2u8]) (:44) and of
// RandomTape::new(b"proof") (VP/commit_test.rs:74).
static inline FlowResult run_flow(const BuiltInstance &bi, const Fl &seed_q, const Fl &seed_p, bool do_verify) {
  FlowResult fr;
  memset(&fr.times, 0, sizeof(fr.times));
  R1CSInstance inst;
  R1CSError e = instance_new(bi.num_cons, bi.num_vars, bi.num_inputs, bi.A.data(), bi.A.size(), bi.B.data(), bi.B.size(),
                             bi.C.data(), bi.C.size(), &inst);
  if (e != R1CS_OK) throw std::runtime_error("instance_new failed");
  FlVec vars_para, vars_input, vars, inputs;
  if (!assignment_new(bi.vars_para, inst.num_vars, &vars_para) || !assignment_new(bi.vars_input, inst.num_vars, &vars_input) ||
      !assignment_new(bi.vars, inst.num_vars, &vars) || !assignment_new(bi.inputs, 0, &inputs))
    throw std::runtime_error("invalid scalar");
  if (!inst.is_sat(vars, inputs)) throw std::runtime_error("instance not satisfied");
  double t0 = now_ms();
  SNARKGens gens = snarkgens_new(bi.num_cons, bi.num_vars, bi.num_inputs, bi.num_non_zero_entries);
  fr.ms_gens = now_ms() - t0;
  t0 = now_ms();
  MultiSparseMatPolynomialAsDense decomm;
  R1CSCommitment comm = snark_encode(inst, gens, &decomm);
  fr.times.ms[T_ENCODE] = now_ms() - t0;
  fr.comm = serialize(comm);
  t0 = now_ms();
  const uint8_t two = 2;
  RandomTape tape1((const char *)&two, 1, seed_q);
  DensePoly poly_para(vars_para), poly_input(vars_input), poly_vars(vars);
  FlVec blind_para, blind_input, blind_vars;
  fr.comm_vars_para = dense_commit(poly_para, gens.gens_r1cs_sat.gens_pc, &tape1, &blind_para);
  fr.comm_vars_input = dense_commit(poly_input, gens.gens_r1cs_sat.gens_pc, &tape1, &blind_input);
  fr.comm_vars = my_dense_mlpoly_commit(poly_vars, gens.gens_r1cs_sat.gens_pc, blind_para, blind_input, &blind_vars);
  PolyCommitment combine;
  for (size_t i = 0; i < fr.comm_vars_para.C.size(); i++)
    combine.C.push_back(compress(pt_add(decompress_or_die(fr.comm_vars_para.C[i]), decompress_or_die(fr.comm_vars_input.C[i]))));
  if (!(combine.C[0] == fr.comm_vars.C[0])) throw std::runtime_error("commitment homomorphism check failed");
  fr.ms_commits = now_ms() - t0;
  Transcript pt("snark_example");
  SNARK proof = my_lib_prove(inst, decomm, vars, inputs, gens, pt, poly_vars, combine, blind_vars, seed_p, &fr.times);
  fr.proof = serialize(proof);
  fr.verified = false;
  fr.ms_verify = 0;
  if (do_verify) {
    t0 = now_ms();
    Transcript vt("snark_example");
    fr.verified = my_lib_verify(proof, comm, inputs, vt, gens, fr.comm_vars_para, fr.comm_vars_input);
    fr.ms_verify = now_ms() - t0;
  }
  return fr;
}

}  // namespace orc
