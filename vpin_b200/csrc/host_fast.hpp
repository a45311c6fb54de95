// Host-only 64-bit arithmetic for the O(1)-size sigma protocols that stay on the CPU next to the Fiat-Shamir transcript
// (Spartan/src/nizk/mod.rs, the per-round commitments of Spartan/src/sumcheck.rs:660-748): F_p (p = 2^255 - 19) in
// 5 x 51-bit limbs with 128-bit products, ristretto255 points on top of it, fixed-base scalar multiplication with signed
// 8-bit windows. The device representation (8 x u32, ed.cuh) converts in and out at the edges. The prover does ~16
// fixed-base multiplications and ~5 encodings per sumcheck round on the host while the GPU runs the next round, so these
// have to cost microseconds, which the portable 32-bit-limb host build of ed.cuh does not deliver.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "ed.cuh"

namespace vpin {
namespace hf {

typedef unsigned __int128 u128;
static const uint64_t kM51 = ((uint64_t)1 << 51) - 1;

// limbs of a "reduced" element are < 2^52; mul / sqr accept limbs < 2^54 and return reduced elements
struct fe { uint64_t v[5]; };

static inline fe fe_zero() { return fe{{0, 0, 0, 0, 0}}; }
static inline fe fe_one() { return fe{{1, 0, 0, 0, 0}}; }
// from the device representation: any value < 2^256 (ed.cuh keeps field elements lazily reduced)
static inline fe fe_from_fp(const fp_t &a) {
  uint64_t x[4];
  for (int i = 0; i < 4; i++) x[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
  fe r;
  r.v[0] = x[0] & kM51;
  r.v[1] = ((x[0] >> 51) | (x[1] << 13)) & kM51;
  r.v[2] = ((x[1] >> 38) | (x[2] << 26)) & kM51;
  r.v[3] = ((x[2] >> 25) | (x[3] << 39)) & kM51;
  r.v[4] = x[3] >> 12;  // 52 bits
  return r;
}
static inline fe fe_carry(fe a) {
  uint64_t c;
  c = a.v[0] >> 51; a.v[0] &= kM51; a.v[1] += c;
  c = a.v[1] >> 51; a.v[1] &= kM51; a.v[2] += c;
  c = a.v[2] >> 51; a.v[2] &= kM51; a.v[3] += c;
  c = a.v[3] >> 51; a.v[3] &= kM51; a.v[4] += c;
  c = a.v[4] >> 51; a.v[4] &= kM51; a.v[0] += 19 * c;
  return a;
}
static inline fe fe_add(const fe &a, const fe &b) {  // reduced + reduced -> limbs < 2^53
  fe r;
  for (int i = 0; i < 5; i++) r.v[i] = a.v[i] + b.v[i];
  return r;
}
// a - b for limbs of a, b < 2^54 - 152: adds 8p, returns a reduced element
static inline fe fe_sub(const fe &a, const fe &b) {
  const uint64_t k0 = ((uint64_t)1 << 54) - 152, k = ((uint64_t)1 << 54) - 8;
  fe r;
  r.v[0] = a.v[0] + k0 - b.v[0];
  for (int i = 1; i < 5; i++) r.v[i] = a.v[i] + k - b.v[i];
  return fe_carry(r);
}
static inline fe fe_neg(const fe &a) { return fe_sub(fe_zero(), a); }
static inline fe fe_reduce_wide(u128 t0, u128 t1, u128 t2, u128 t3, u128 t4) {
  fe r;
  uint64_t c;
  r.v[0] = (uint64_t)t0 & kM51; c = (uint64_t)(t0 >> 51);
  t1 += c; r.v[1] = (uint64_t)t1 & kM51; c = (uint64_t)(t1 >> 51);
  t2 += c; r.v[2] = (uint64_t)t2 & kM51; c = (uint64_t)(t2 >> 51);
  t3 += c; r.v[3] = (uint64_t)t3 & kM51; c = (uint64_t)(t3 >> 51);
  t4 += c; r.v[4] = (uint64_t)t4 & kM51; c = (uint64_t)(t4 >> 51);
  u128 z = (u128)c * 19 + r.v[0];
  r.v[0] = (uint64_t)z & kM51;
  r.v[1] += (uint64_t)(z >> 51);
  return r;
}
static inline fe fe_mul(const fe &a, const fe &b) {
  uint64_t r0 = a.v[0], r1 = a.v[1], r2 = a.v[2], r3 = a.v[3], r4 = a.v[4];
  uint64_t s0 = b.v[0], s1 = b.v[1], s2 = b.v[2], s3 = b.v[3], s4 = b.v[4];
  u128 t0 = (u128)r0 * s0;
  u128 t1 = (u128)r0 * s1 + (u128)r1 * s0;
  u128 t2 = (u128)r0 * s2 + (u128)r2 * s0 + (u128)r1 * s1;
  u128 t3 = (u128)r0 * s3 + (u128)r3 * s0 + (u128)r1 * s2 + (u128)r2 * s1;
  u128 t4 = (u128)r0 * s4 + (u128)r4 * s0 + (u128)r3 * s1 + (u128)r1 * s3 + (u128)r2 * s2;
  r1 *= 19; r2 *= 19; r3 *= 19; r4 *= 19;
  t0 += (u128)r4 * s1 + (u128)r1 * s4 + (u128)r2 * s3 + (u128)r3 * s2;
  t1 += (u128)r4 * s2 + (u128)r2 * s4 + (u128)r3 * s3;
  t2 += (u128)r4 * s3 + (u128)r3 * s4;
  t3 += (u128)r4 * s4;
  return fe_reduce_wide(t0, t1, t2, t3, t4);
}
static inline fe fe_sqr(const fe &a) {
  uint64_t r0 = a.v[0], r1 = a.v[1], r2 = a.v[2], r3 = a.v[3], r4 = a.v[4];
  uint64_t d0 = r0 * 2, d1 = r1 * 2, d2 = r2 * 2 * 19, d419 = r4 * 19, d4 = d419 * 2;
  u128 t0 = (u128)r0 * r0 + (u128)d4 * r1 + (u128)d2 * r3;
  u128 t1 = (u128)d0 * r1 + (u128)d4 * r2 + (u128)r3 * (r3 * 19);
  u128 t2 = (u128)d0 * r2 + (u128)r1 * r1 + (u128)d4 * r3;
  u128 t3 = (u128)d0 * r3 + (u128)d1 * r2 + (u128)r4 * d419;
  u128 t4 = (u128)d0 * r4 + (u128)d1 * r3 + (u128)r2 * r2;
  return fe_reduce_wide(t0, t1, t2, t3, t4);
}
static inline fe fe_sqr_n(fe a, int n) { for (int i = 0; i < n; i++) a = fe_sqr(a); return a; }
// canonical little-endian bytes
static inline void fe_to_bytes(const fe &a, uint8_t out[32]) {
  fe t = fe_carry(fe_carry(a));
  // t < 2^255 + small: add 19, the carry out of bit 255 tells whether t >= p
  uint64_t q = (t.v[0] + 19) >> 51;
  q = (t.v[1] + q) >> 51;
  q = (t.v[2] + q) >> 51;
  q = (t.v[3] + q) >> 51;
  q = (t.v[4] + q) >> 51;
  t.v[0] += 19 * q;
  uint64_t c;
  c = t.v[0] >> 51; t.v[0] &= kM51; t.v[1] += c;
  c = t.v[1] >> 51; t.v[1] &= kM51; t.v[2] += c;
  c = t.v[2] >> 51; t.v[2] &= kM51; t.v[3] += c;
  c = t.v[3] >> 51; t.v[3] &= kM51; t.v[4] += c;
  t.v[4] &= kM51;
  uint64_t w[4];
  w[0] = t.v[0] | (t.v[1] << 51);
  w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
  w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
  w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
  memcpy(out, w, 32);
}
static inline bool fe_is_neg(const fe &a) { uint8_t b[32]; fe_to_bytes(a, b); return b[0] & 1; }
static inline bool fe_is_zero(const fe &a) {
  uint8_t b[32];
  fe_to_bytes(a, b);
  uint8_t o = 0;
  for (int i = 0; i < 32; i++) o |= b[i];
  return o == 0;
}
static inline bool fe_eq(const fe &a, const fe &b) { return fe_is_zero(fe_sub(a, b)); }
static inline fe fe_abs(const fe &a) { return fe_is_neg(a) ? fe_neg(a) : a; }
// a^(2^252 - 3)
static inline fe fe_pow22523(const fe &z) {
  fe t0 = fe_sqr(z);
  fe t1 = fe_mul(z, fe_sqr_n(t0, 2));
  t0 = fe_mul(t0, t1);
  t0 = fe_mul(t1, fe_sqr(t0));
  t1 = fe_mul(fe_sqr_n(t0, 5), t0);
  fe t2 = fe_mul(fe_sqr_n(t1, 10), t1);
  fe t3 = fe_mul(fe_sqr_n(t2, 20), t2);
  t2 = fe_mul(fe_sqr_n(t3, 10), t1);
  t3 = fe_mul(fe_sqr_n(t2, 50), t2);
  fe t4 = fe_mul(fe_sqr_n(t3, 100), t3);
  t3 = fe_mul(fe_sqr_n(t4, 50), t2);
  return fe_mul(fe_sqr_n(t3, 2), z);
}
static inline fe fe_invert(const fe &z) {
  fe t = fe_sqr_n(fe_pow22523(z), 3);
  return fe_mul(t, fe_mul(fe_sqr(z), z));
}
struct Consts { fe d2, sqrt_m1, invsqrt_a_minus_d; };
static inline const Consts &consts() {
  static const Consts c = {fe_from_fp(fp_d2()), fe_from_fp(fp_sqrt_m1()), fe_from_fp(fp_invsqrt_a_minus_d())};
  return c;
}
// RFC 9496 4.2 SQRT_RATIO_M1
static inline bool fe_sqrt_ratio_m1(const fe &u, const fe &v, fe *out) {
  fe v3 = fe_mul(fe_sqr(v), v);
  fe v7 = fe_mul(fe_sqr(v3), v);
  fe r = fe_mul(fe_mul(u, v3), fe_pow22523(fe_mul(u, v7)));
  fe check = fe_mul(v, fe_sqr(r));
  fe neg_u = fe_neg(u);
  bool correct = fe_eq(check, u);
  bool flipped = fe_eq(check, neg_u);
  bool flipped_i = fe_eq(check, fe_mul(neg_u, consts().sqrt_m1));
  if (flipped || flipped_i) r = fe_mul(r, consts().sqrt_m1);
  *out = fe_abs(r);
  return correct || flipped;
}

// ---- points (all coordinates reduced) ----
struct ge { fe X, Y, Z, T; };
struct niels { fe yp, ym, t2d; };
static inline ge ge_identity() { return ge{fe_zero(), fe_one(), fe_one(), fe_zero()}; }
static inline ge ge_from_dev(const ge_t &p) { return ge{fe_from_fp(p.X), fe_from_fp(p.Y), fe_from_fp(p.Z), fe_from_fp(p.T)}; }
static inline ge ge_finish(const fe &e, const fe &f, const fe &g, const fe &h) {
  return ge{fe_mul(e, f), fe_mul(g, h), fe_mul(f, g), fe_mul(e, h)};
}
static inline ge ge_madd(const ge &p, const niels &q) {
  fe a = fe_mul(fe_sub(p.Y, p.X), q.ym);
  fe b = fe_mul(fe_add(p.Y, p.X), q.yp);
  fe c = fe_mul(p.T, q.t2d);
  fe d = fe_add(p.Z, p.Z);
  return ge_finish(fe_sub(b, a), fe_sub(d, c), fe_carry(fe_add(d, c)), fe_add(b, a));
}
static inline ge ge_msub(const ge &p, const niels &q) {
  fe a = fe_mul(fe_sub(p.Y, p.X), q.yp);
  fe b = fe_mul(fe_add(p.Y, p.X), q.ym);
  fe c = fe_mul(p.T, q.t2d);
  fe d = fe_add(p.Z, p.Z);
  return ge_finish(fe_sub(b, a), fe_carry(fe_add(d, c)), fe_sub(d, c), fe_add(b, a));
}
static inline ge ge_add(const ge &p, const ge &q) {
  fe a = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
  fe b = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
  fe c = fe_mul(fe_mul(p.T, q.T), consts().d2);
  fe d = fe_mul(p.Z, q.Z);
  d = fe_add(d, d);
  return ge_finish(fe_sub(b, a), fe_sub(d, c), fe_carry(fe_add(d, c)), fe_add(b, a));
}
static inline ge ge_dbl(const ge &p) {
  fe a = fe_sqr(p.X), b = fe_sqr(p.Y), c = fe_sqr(p.Z);
  c = fe_add(c, c);
  fe xy = fe_add(p.X, p.Y);
  fe e = fe_sub(fe_sqr(xy), fe_add(a, b));
  fe g = fe_sub(b, a), h = fe_neg(fe_carry(fe_add(a, b)));  // g = -a + b, h = -a - b
  fe f = fe_sub(g, c);
  return ge_finish(e, f, g, h);
}
static inline niels ge_to_niels(const ge &p, const fe &zinv) {
  fe x = fe_mul(p.X, zinv), y = fe_mul(p.Y, zinv);
  return niels{fe_carry(fe_add(y, x)), fe_sub(y, x), fe_mul(fe_mul(x, y), consts().d2)};
}
// RFC 9496 4.3.2 ENCODE (dalek RistrettoPoint::compress)
static inline void ge_compress(const ge &p, uint8_t out[32]) {
  const Consts &k = consts();
  fe u1 = fe_mul(fe_add(p.Z, p.Y), fe_sub(p.Z, p.Y));
  fe u2 = fe_mul(p.X, p.Y);
  fe invsqrt;
  fe_sqrt_ratio_m1(fe_one(), fe_mul(u1, fe_sqr(u2)), &invsqrt);
  fe den1 = fe_mul(invsqrt, u1), den2 = fe_mul(invsqrt, u2);
  fe z_inv = fe_mul(fe_mul(den1, den2), p.T);
  fe ix0 = fe_mul(p.X, k.sqrt_m1), iy0 = fe_mul(p.Y, k.sqrt_m1);
  fe enchanted = fe_mul(den1, k.invsqrt_a_minus_d);
  bool rotate = fe_is_neg(fe_mul(p.T, z_inv));
  fe x = rotate ? iy0 : p.X, y = rotate ? ix0 : p.Y, den_inv = rotate ? enchanted : den2;
  if (fe_is_neg(fe_mul(x, z_inv))) y = fe_neg(y);
  fe s = fe_abs(fe_mul(den_inv, fe_sub(p.Z, y)));
  fe_to_bytes(s, out);
}

// Fixed-base scalar multiplication: 32 positions x 128 multiples (signed 8-bit digits), affine Niels entries.
struct FixedBase {
  std::vector<niels> tbl;
  void build(const ge &p) {
    const int P = 32, M = 128;
    std::vector<ge> ext((size_t)P * M);
    ge base = p;
    for (int pos = 0; pos < P; pos++) {
      ge cur = base;
      for (int m = 0; m < M; m++) {
        ext[(size_t)pos * M + m] = cur;
        if (m + 1 < M) cur = ge_add(cur, base);
      }
      for (int k = 0; k < 8; k++) base = ge_dbl(base);
    }
    std::vector<fe> prefix((size_t)P * M);
    fe run = fe_one();
    for (int i = 0; i < P * M; i++) { run = fe_mul(run, ext[i].Z); prefix[i] = run; }
    fe inv = fe_invert(run);
    tbl.resize((size_t)P * M);
    for (int i = P * M - 1; i >= 0; i--) {
      fe zinv = i > 0 ? fe_mul(inv, prefix[i - 1]) : inv;
      inv = fe_mul(inv, ext[i].Z);
      tbl[i] = ge_to_niels(ext[i], zinv);
    }
  }
  // *acc += s * P, s in Montgomery form
  void mul_acc(const fl_t &s_mont, ge *acc) const {
    fl_t s = fl_from_mont(s_mont);
    const uint8_t *b = reinterpret_cast<const uint8_t *>(s.v);
    int carry = 0;
    for (int pos = 0; pos < 32; pos++) {
      int d = (int)b[pos] + carry;
      carry = 0;
      if (d > 128) { d -= 256; carry = 1; }
      if (d > 0) *acc = ge_madd(*acc, tbl[(size_t)pos * 128 + d - 1]);
      else if (d < 0) *acc = ge_msub(*acc, tbl[(size_t)pos * 128 - d - 1]);
    }
  }
  ge mul(const fl_t &s_mont) const {
    ge acc = ge_identity();
    mul_acc(s_mont, &acc);
    return acc;
  }
};

}  // namespace hf
}  // namespace vpin
