// TEST INFRASTRUCTURE ONLY (CPU oracle). Not linked into the product library.
//
// ristretto255 group used by Spartan as GroupElement (Spartan/src/group.rs:6-8). The arithmetic lives in
// curve25519-dalek 3.2.0, which is NOT vendored under /root/reference (Spartan/Cargo.toml:14,
// vPIN_proof_generation/Cargo.lock). This file restates the published algorithms:
//   * F_p, p = 2^255-19, radix-2^51 (the layout of dalek's u64 backend / curve25519-donna)
//   * extended twisted Edwards a=-1 addition/doubling (Hisil-Wong-Carter-Dawson 2008)
//   * RFC 9496 ristretto255 DECODE / ENCODE / MAP (== dalek compress/decompress/from_uniform_bytes)
//   * vartime multiscalar mul: Straus width-5 NAF below 190 points, Pippenger above — dalek 3.2.0's policy
//     (used only through Spartan/src/group.rs:103-121). Results are algorithm-independent after ENCODE.
// Pinned by tests/test_oracle_primitives.py against RFC 9496 vectors, libsodium 1.0.20 and oracle/pyref.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "fl.hpp"

namespace orc {

struct Fp { uint64_t v[5]; };
static const uint64_t M51 = (1ULL << 51) - 1;

static inline Fp fp_zero() { return Fp{{0, 0, 0, 0, 0}}; }
static inline Fp fp_one() { return Fp{{1, 0, 0, 0, 0}}; }
static inline void fp_carry(Fp &a) {
  uint64_t c;
  c = a.v[0] >> 51; a.v[0] &= M51; a.v[1] += c;
  c = a.v[1] >> 51; a.v[1] &= M51; a.v[2] += c;
  c = a.v[2] >> 51; a.v[2] &= M51; a.v[3] += c;
  c = a.v[3] >> 51; a.v[3] &= M51; a.v[4] += c;
  c = a.v[4] >> 51; a.v[4] &= M51; a.v[0] += c * 19;
}
static inline Fp fp_add(const Fp &a, const Fp &b) {
  Fp r;
  for (int i = 0; i < 5; i++) r.v[i] = a.v[i] + b.v[i];
  fp_carry(r);
  return r;
}
static inline Fp fp_sub(const Fp &a, const Fp &b) {
  // add 16p limb-wise so no limb underflows (inputs are < 2^52)
  Fp r;
  r.v[0] = a.v[0] + 36028797018963664ULL - b.v[0];
  for (int i = 1; i < 5; i++) r.v[i] = a.v[i] + 36028797018963952ULL - b.v[i];
  fp_carry(r);
  return r;
}
static inline Fp fp_neg(const Fp &a) { return fp_sub(fp_zero(), a); }
static inline Fp fp_mul(const Fp &a, const Fp &b) {
  const uint64_t *x = a.v, *y = b.v;
  uint64_t y1_19 = y[1] * 19, y2_19 = y[2] * 19, y3_19 = y[3] * 19, y4_19 = y[4] * 19;
  u128 c0 = (u128)x[0] * y[0] + (u128)x[4] * y1_19 + (u128)x[3] * y2_19 + (u128)x[2] * y3_19 + (u128)x[1] * y4_19;
  u128 c1 = (u128)x[1] * y[0] + (u128)x[0] * y[1] + (u128)x[4] * y2_19 + (u128)x[3] * y3_19 + (u128)x[2] * y4_19;
  u128 c2 = (u128)x[2] * y[0] + (u128)x[1] * y[1] + (u128)x[0] * y[2] + (u128)x[4] * y3_19 + (u128)x[3] * y4_19;
  u128 c3 = (u128)x[3] * y[0] + (u128)x[2] * y[1] + (u128)x[1] * y[2] + (u128)x[0] * y[3] + (u128)x[4] * y4_19;
  u128 c4 = (u128)x[4] * y[0] + (u128)x[3] * y[1] + (u128)x[2] * y[2] + (u128)x[1] * y[3] + (u128)x[0] * y[4];
  Fp r;
  c1 += (uint64_t)(c0 >> 51); r.v[0] = (uint64_t)c0 & M51;
  c2 += (uint64_t)(c1 >> 51); r.v[1] = (uint64_t)c1 & M51;
  c3 += (uint64_t)(c2 >> 51); r.v[2] = (uint64_t)c2 & M51;
  c4 += (uint64_t)(c3 >> 51); r.v[3] = (uint64_t)c3 & M51;
  uint64_t carry = (uint64_t)(c4 >> 51); r.v[4] = (uint64_t)c4 & M51;
  r.v[0] += carry * 19;
  r.v[1] += r.v[0] >> 51; r.v[0] &= M51;
  return r;
}
static inline Fp fp_sqr(const Fp &a) { return fp_mul(a, a); }
static inline Fp fp_mul_small(const Fp &a, uint64_t k) { Fp b = {{k, 0, 0, 0, 0}}; return fp_mul(a, b); }

static inline void fp_tobytes(const Fp &a, uint8_t out[32]) {
  Fp t = a;
  fp_carry(t);
  fp_carry(t);
  // now t < 2^255 + small; compute q = (t + 19) >> 255
  uint64_t q = (t.v[0] + 19) >> 51;
  q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
  t.v[0] += 19 * q;
  uint64_t c;
  c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
  c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
  c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
  c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
  t.v[4] &= M51;
  uint64_t w[4];
  w[0] = t.v[0] | (t.v[1] << 51);
  w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
  w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
  w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
  memcpy(out, w, 32);
}
static inline Fp fp_frombytes(const uint8_t in[32]) {  // ignores bit 255
  uint64_t w[4];
  memcpy(w, in, 32);
  Fp r;
  r.v[0] = w[0] & M51;
  r.v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
  r.v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
  r.v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
  r.v[4] = (w[3] >> 12) & M51;
  return r;
}
static inline bool fp_eq(const Fp &a, const Fp &b) {
  uint8_t x[32], y[32];
  fp_tobytes(a, x);
  fp_tobytes(b, y);
  return memcmp(x, y, 32) == 0;
}
static inline bool fp_is_neg(const Fp &a) { uint8_t x[32]; fp_tobytes(a, x); return x[0] & 1; }
static inline bool fp_is_zero(const Fp &a) { return fp_eq(a, fp_zero()); }
static inline Fp fp_abs(const Fp &a) { return fp_is_neg(a) ? fp_neg(a) : a; }
static inline Fp fp_pow(const Fp &a, const uint64_t e[4]) {
  Fp r = fp_one();
  for (int i = 3; i >= 0; i--)
    for (int j = 63; j >= 0; j--) {
      r = fp_sqr(r);
      if ((e[i] >> j) & 1) r = fp_mul(r, a);
    }
  return r;
}
static inline Fp fp_invert(const Fp &a) {  // a^(p-2)
  static const uint64_t e[4] = {0xffffffffffffffebULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0x7fffffffffffffffULL};
  return fp_pow(a, e);
}
static inline Fp fp_pow_p58(const Fp &a) {  // a^((p-5)/8) = a^(2^252-3)
  static const uint64_t e[4] = {0xfffffffffffffffdULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0x0fffffffffffffffULL};
  return fp_pow(a, e);
}
static inline Fp fp_from_hex_le(const char *hex) {  // 64 hex chars, little-endian bytes
  uint8_t b[32];
  for (int i = 0; i < 32; i++) {
    auto nib = [](char c) -> int { return c <= '9' ? c - '0' : c - 'a' + 10; };
    b[i] = (uint8_t)(nib(hex[2 * i]) * 16 + nib(hex[2 * i + 1]));
  }
  return fp_frombytes(b);
}

// RFC 9496 section 4.1 constants, little-endian byte strings (values recomputed in tests/test_oracle_primitives.py)
struct EdConsts {
  Fp d, d2, sqrt_m1, sqrt_ad_minus_one, invsqrt_a_minus_d, one_minus_d_sq, d_minus_one_sq;
  EdConsts() {
    d = fp_from_hex_le("a3785913ca4deb75abd841414d0a700098e879777940c78c73fe6f2bee6c0352");
    d2 = fp_add(d, d);
    sqrt_m1 = fp_from_hex_le("b0a00e4a271beec478e42fad0618432fa7d7fb3d99004d2b0bdfc14f8024832b");
    sqrt_ad_minus_one = fp_from_hex_le("1b2e7b49a0f6977ebd54781b0c8e9daffdd1f531c9fc3c0fac48832bbf316937");
    invsqrt_a_minus_d = fp_from_hex_le("ea405d80aafdc899be72415a17162f9d40d801fe917bc216a2fcafcf05896c78");
    Fp one = fp_one();
    one_minus_d_sq = fp_sub(one, fp_sqr(d));
    Fp dm1 = fp_sub(d, one);
    d_minus_one_sq = fp_sqr(dm1);
  }
};
static inline const EdConsts &edc() { static EdConsts c; return c; }

// RFC 9496 4.2 SQRT_RATIO_M1
static inline bool fp_sqrt_ratio_m1(const Fp &u, const Fp &v, Fp *out) {
  const EdConsts &C = edc();
  Fp v3 = fp_mul(fp_sqr(v), v);
  Fp v7 = fp_mul(fp_sqr(v3), v);
  Fp r = fp_mul(fp_mul(u, v3), fp_pow_p58(fp_mul(u, v7)));
  Fp check = fp_mul(v, fp_sqr(r));
  Fp neg_u = fp_neg(u);
  bool correct = fp_eq(check, u);
  bool flipped = fp_eq(check, neg_u);
  bool flipped_i = fp_eq(check, fp_mul(neg_u, C.sqrt_m1));
  if (flipped || flipped_i) r = fp_mul(r, C.sqrt_m1);
  *out = fp_abs(r);
  return correct || flipped;
}

struct Pt { Fp X, Y, Z, T; };
struct PtCached { Fp YpX, YmX, Z, T2d; };

static inline Pt pt_identity() { return Pt{fp_zero(), fp_one(), fp_one(), fp_zero()}; }
static inline Pt pt_neg(const Pt &p) { return Pt{fp_neg(p.X), p.Y, p.Z, fp_neg(p.T)}; }
static inline PtCached pt_to_cached(const Pt &p) {
  return PtCached{fp_add(p.Y, p.X), fp_sub(p.Y, p.X), p.Z, fp_mul(p.T, edc().d2)};
}
static inline PtCached cached_neg(const PtCached &c) { return PtCached{c.YmX, c.YpX, c.Z, fp_neg(c.T2d)}; }
static inline Pt pt_add_cached(const Pt &p, const PtCached &q) {
  Fp a = fp_mul(fp_sub(p.Y, p.X), q.YmX);
  Fp b = fp_mul(fp_add(p.Y, p.X), q.YpX);
  Fp c = fp_mul(p.T, q.T2d);
  Fp d = fp_mul(p.Z, q.Z);
  d = fp_add(d, d);
  Fp e = fp_sub(b, a), f = fp_sub(d, c), g = fp_add(d, c), h = fp_add(b, a);
  return Pt{fp_mul(e, f), fp_mul(g, h), fp_mul(f, g), fp_mul(e, h)};
}
static inline Pt pt_add(const Pt &p, const Pt &q) { return pt_add_cached(p, pt_to_cached(q)); }
static inline Pt pt_sub(const Pt &p, const Pt &q) { return pt_add(p, pt_neg(q)); }
static inline Pt pt_double(const Pt &p) {
  Fp a = fp_sqr(p.X), b = fp_sqr(p.Y), c = fp_sqr(p.Z);
  c = fp_add(c, c);
  Fp d = fp_neg(a);
  Fp xy = fp_add(p.X, p.Y);
  Fp e = fp_sub(fp_sub(fp_sqr(xy), a), b);
  Fp g = fp_add(d, b), f = fp_sub(g, c), h = fp_sub(d, b);
  return Pt{fp_mul(e, f), fp_mul(g, h), fp_mul(f, g), fp_mul(e, h)};
}
// ristretto equality: X1*Y2 == Y1*X2 or Y1*Y2 == X1*X2
static inline bool pt_eq(const Pt &p, const Pt &q) {
  return fp_eq(fp_mul(p.X, q.Y), fp_mul(p.Y, q.X)) || fp_eq(fp_mul(p.Y, q.Y), fp_mul(p.X, q.X));
}

// RFC 9496 4.3.2 ENCODE
static inline void pt_compress(const Pt &p, uint8_t out[32]) {
  const EdConsts &C = edc();
  Fp u1 = fp_mul(fp_add(p.Z, p.Y), fp_sub(p.Z, p.Y));
  Fp u2 = fp_mul(p.X, p.Y);
  Fp invsqrt;
  fp_sqrt_ratio_m1(fp_one(), fp_mul(u1, fp_sqr(u2)), &invsqrt);
  Fp den1 = fp_mul(invsqrt, u1), den2 = fp_mul(invsqrt, u2);
  Fp z_inv = fp_mul(fp_mul(den1, den2), p.T);
  Fp ix0 = fp_mul(p.X, C.sqrt_m1), iy0 = fp_mul(p.Y, C.sqrt_m1);
  Fp enchanted = fp_mul(den1, C.invsqrt_a_minus_d);
  bool rotate = fp_is_neg(fp_mul(p.T, z_inv));
  Fp x = rotate ? iy0 : p.X, y = rotate ? ix0 : p.Y, den_inv = rotate ? enchanted : den2;
  if (fp_is_neg(fp_mul(x, z_inv))) y = fp_neg(y);
  Fp s = fp_abs(fp_mul(den_inv, fp_sub(p.Z, y)));
  fp_tobytes(s, out);
}
// RFC 9496 4.3.1 DECODE
static inline bool pt_decompress(const uint8_t in[32], Pt *out) {
  const EdConsts &C = edc();
  Fp s = fp_frombytes(in);
  uint8_t chk[32];
  fp_tobytes(s, chk);
  if (memcmp(chk, in, 32) != 0 || (in[0] & 1)) return false;  // non-canonical or negative
  Fp one = fp_one();
  Fp ss = fp_sqr(s);
  Fp u1 = fp_sub(one, ss), u2 = fp_add(one, ss);
  Fp u2_sqr = fp_sqr(u2);
  Fp v = fp_sub(fp_neg(fp_mul(C.d, fp_sqr(u1))), u2_sqr);
  Fp invsqrt;
  bool ok = fp_sqrt_ratio_m1(one, fp_mul(v, u2_sqr), &invsqrt);
  Fp den_x = fp_mul(invsqrt, u2);
  Fp den_y = fp_mul(fp_mul(invsqrt, den_x), v);
  Fp x = fp_abs(fp_mul(fp_add(s, s), den_x));
  Fp y = fp_mul(u1, den_y);
  Fp t = fp_mul(x, y);
  if (!ok || fp_is_neg(t) || fp_is_zero(y)) return false;
  *out = Pt{x, y, one, t};
  return true;
}
// RFC 9496 4.3.4 MAP
static inline Pt pt_elligator(const Fp &t) {
  const EdConsts &C = edc();
  Fp one = fp_one();
  Fp r = fp_mul(C.sqrt_m1, fp_sqr(t));
  Fp u = fp_mul(fp_add(r, one), C.one_minus_d_sq);
  Fp v = fp_mul(fp_sub(fp_neg(one), fp_mul(r, C.d)), fp_add(r, C.d));
  Fp s;
  bool was_square = fp_sqrt_ratio_m1(u, v, &s);
  Fp s_prime = fp_neg(fp_abs(fp_mul(s, t)));
  if (!was_square) s = s_prime;
  Fp c = was_square ? fp_neg(one) : r;
  Fp n = fp_sub(fp_mul(fp_mul(c, fp_sub(r, one)), C.d_minus_one_sq), v);
  Fp w0 = fp_mul(fp_add(s, s), v);
  Fp w1 = fp_mul(n, C.sqrt_ad_minus_one);
  Fp ss = fp_sqr(s);
  Fp w2 = fp_sub(one, ss), w3 = fp_add(one, ss);
  return Pt{fp_mul(w0, w3), fp_mul(w2, w1), fp_mul(w1, w3), fp_mul(w0, w2)};
}
// dalek RistrettoPoint::from_uniform_bytes (used at Spartan/src/commitments.rs:30)
static inline Pt pt_from_uniform_bytes(const uint8_t b[64]) {
  Fp t1 = fp_frombytes(b), t2 = fp_frombytes(b + 32);
  return pt_add(pt_elligator(t1), pt_elligator(t2));
}
static const uint8_t BASEPOINT_COMPRESSED[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9,
                                                 0x61, 0xc5, 0x00, 0x51, 0x5f, 0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82,
                                                 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};

// ---------------------------------------------------------------------------------------------------
// vartime multiscalar multiplication (Spartan/src/group.rs:103-121 -> dalek). Scalars arrive as F_l
// Montgomery values and are converted with to_bytes (Spartan/src/scalar/mod.rs:38-40).

static inline void scalar_naf5(const uint8_t s[32], int8_t naf[256]) {
  // width-5 non-adjacent form
  memset(naf, 0, 256);
  uint64_t x[5] = {0, 0, 0, 0, 0};
  memcpy(x, s, 32);
  const int w = 5;
  const uint64_t width = 1 << w, window_mask = width - 1;
  int pos = 0;
  uint64_t carry = 0;
  while (pos < 256) {
    int idx = pos / 64, bit = pos % 64;
    uint64_t bit_buf = bit < 64 - w ? x[idx] >> bit : (x[idx] >> bit) | (x[idx + 1] << (64 - bit));
    uint64_t window = carry + (bit_buf & window_mask);
    if ((window & 1) == 0) { pos += 1; continue; }
    if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
    else { carry = 1; naf[pos] = (int8_t)((int64_t)window - (int64_t)width); }
    pos += w;
  }
}

static inline Pt msm_straus(const std::vector<const uint8_t *> &scalars, const Pt *points, size_t n) {
  std::vector<int8_t> nafs(n * 256);
  std::vector<PtCached> tables(n * 8);  // odd multiples 1,3,...,15
  for (size_t i = 0; i < n; i++) {
    scalar_naf5(scalars[i], &nafs[i * 256]);
    Pt p2 = pt_double(points[i]);
    Pt cur = points[i];
    tables[i * 8] = pt_to_cached(cur);
    for (int k = 1; k < 8; k++) { cur = pt_add(cur, p2); tables[i * 8 + k] = pt_to_cached(cur); }
  }
  Pt acc = pt_identity();
  bool started = false;
  for (int pos = 255; pos >= 0; pos--) {
    if (started) acc = pt_double(acc);
    for (size_t i = 0; i < n; i++) {
      int8_t d = nafs[i * 256 + pos];
      if (d > 0) { acc = pt_add_cached(acc, tables[i * 8 + d / 2]); started = true; }
      else if (d < 0) { acc = pt_add_cached(acc, cached_neg(tables[i * 8 + (-d) / 2])); started = true; }
    }
  }
  return acc;
}

static inline Pt msm_pippenger(const std::vector<const uint8_t *> &scalars, const Pt *points, size_t n) {
  int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
  int digits_count = (256 + w - 1) / w;
  size_t buckets_count = (size_t)1 << (w - 1);
  // signed radix-2^w digits
  std::vector<int16_t> digits(n * digits_count);
  for (size_t i = 0; i < n; i++) {
    uint64_t x[5] = {0, 0, 0, 0, 0};
    memcpy(x, scalars[i], 32);
    uint64_t radix = 1ULL << w, mask = radix - 1, carry = 0;
    for (int k = 0; k < digits_count; k++) {
      int bit_offset = k * w, idx = bit_offset / 64, bit = bit_offset % 64;
      uint64_t bit_buf = x[idx] >> bit;
      if (bit + w > 64) bit_buf |= x[idx + 1] << (64 - bit);  // x[4] == 0 pads the top
      uint64_t coef = carry + (bit_buf & mask);
      carry = (coef + radix / 2) >> w;
      digits[i * digits_count + k] = (int16_t)((int64_t)coef - (int64_t)(carry << w));
    }
    // scalars are < 2^253 so the recoding never carries out of the top digit
  }
  std::vector<PtCached> cached(n);
  for (size_t i = 0; i < n; i++) cached[i] = pt_to_cached(points[i]);
  Pt total = pt_identity();
  std::vector<Pt> buckets(buckets_count);
  for (int k = digits_count - 1; k >= 0; k--) {
    for (size_t b = 0; b < buckets_count; b++) buckets[b] = pt_identity();
    for (size_t i = 0; i < n; i++) {
      int d = digits[i * digits_count + k];
      if (d > 0) buckets[d - 1] = pt_add_cached(buckets[d - 1], cached[i]);
      else if (d < 0) buckets[-d - 1] = pt_add_cached(buckets[-d - 1], cached_neg(cached[i]));
    }
    Pt interm = buckets[buckets_count - 1], sum = buckets[buckets_count - 1];
    for (size_t b = buckets_count - 1; b-- > 0;) { interm = pt_add(interm, buckets[b]); sum = pt_add(sum, interm); }
    for (int j = 0; j < w; j++) total = pt_double(total);
    total = pt_add(total, sum);
  }
  return total;
}

static inline Pt msm(const Fl *scalars, const Pt *points, size_t n) {
  if (n == 0) return pt_identity();
  std::vector<uint8_t> bytes(n * 32);
  std::vector<const uint8_t *> ptrs(n);
  for (size_t i = 0; i < n; i++) { fl_to_bytes(scalars[i], &bytes[i * 32]); ptrs[i] = &bytes[i * 32]; }
  return n < 190 ? msm_straus(ptrs, points, n) : msm_pippenger(ptrs, points, n);
}
static inline Pt pt_mul(const Fl &s, const Pt &p) { return msm(&s, &p, 1); }

}  // namespace orc
