// TEST INFRASTRUCTURE ONLY (CPU oracle). Not linked into the product library.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// F_l, l = 2^252 + 27742317777372353535851937790883648493, 4 x u64 Montgomery form (R = 2^256).
// Restates Spartan/src/scalar/ristretto255.rs:
//   Scalar            :200      constants MODULUS :249, INV :305, R :308, R2 :316, R3 :324
//   from_bytes        :398-424  to_bytes :426-440  from_bytes_wide :442-473
//   square/mul        :483, :702-726   montgomery_reduce :654-698
//   add :748  sub :730  neg :761  invert :548-602  batch_invert :604-651
// Serde of Scalar is the raw Montgomery limbs (:199-200); every op returns the canonical
// representative in [0, l), so any correct implementation yields identical bytes.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;

struct Fl {
  uint64_t v[4];
  bool operator==(const Fl &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
  bool operator!=(const Fl &o) const { return !(*this == o); }
};

static const uint64_t FL_MOD[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0000000000000000ULL, 0x1000000000000000ULL};
static const uint64_t FL_INV = 0xd2b51da312547e1bULL;  // -(l^{-1}) mod 2^64
static const Fl FL_R = {{0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL}};
static const Fl FL_R2 = {{0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL}};
static const Fl FL_R3 = {{0x2a9e49687b83a2dbULL, 0x278324e6aef7f3ecULL, 0x8065dc6c04ec5b65ULL, 0x0e530b773599cec7ULL}};

static inline Fl fl_zero() { return Fl{{0, 0, 0, 0}}; }
static inline Fl fl_one() { return FL_R; }

static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t &carry) {
  u128 t = (u128)a + b + carry;
  carry = (uint64_t)(t >> 64);
  return (uint64_t)t;
}
static inline uint64_t sbb(uint64_t a, uint64_t b, uint64_t &borrow) {
  u128 t = (u128)a - b - borrow;
  borrow = (uint64_t)(t >> 64) & 1;
  return (uint64_t)t;
}
static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t &carry) {
  u128 t = (u128)a + (u128)b * c + carry;
  carry = (uint64_t)(t >> 64);
  return (uint64_t)t;
}

// ristretto255.rs:730-745
static inline Fl fl_sub(const Fl &a, const Fl &b) {
  uint64_t br = 0;
  Fl d;
  d.v[0] = sbb(a.v[0], b.v[0], br);
  d.v[1] = sbb(a.v[1], b.v[1], br);
  d.v[2] = sbb(a.v[2], b.v[2], br);
  d.v[3] = sbb(a.v[3], b.v[3], br);
  uint64_t mask = 0 - br, c = 0;
  d.v[0] = adc(d.v[0], FL_MOD[0] & mask, c);
  d.v[1] = adc(d.v[1], FL_MOD[1] & mask, c);
  d.v[2] = adc(d.v[2], FL_MOD[2] & mask, c);
  d.v[3] = adc(d.v[3], FL_MOD[3] & mask, c);
  return d;
}
// ristretto255.rs:748-758
static inline Fl fl_add(const Fl &a, const Fl &b) {
  uint64_t c = 0;
  Fl d;
  d.v[0] = adc(a.v[0], b.v[0], c);
  d.v[1] = adc(a.v[1], b.v[1], c);
  d.v[2] = adc(a.v[2], b.v[2], c);
  d.v[3] = adc(a.v[3], b.v[3], c);
  Fl m = {{FL_MOD[0], FL_MOD[1], FL_MOD[2], FL_MOD[3]}};
  return fl_sub(d, m);
}
static inline Fl fl_neg(const Fl &a) { return fl_sub(fl_zero(), a); }

// ristretto255.rs:654-698
static inline Fl fl_mont_reduce(uint64_t r0, uint64_t r1, uint64_t r2, uint64_t r3, uint64_t r4, uint64_t r5,
                                uint64_t r6, uint64_t r7) {
  uint64_t k, carry, carry2;
  k = r0 * FL_INV; carry = 0;
  (void)mac(r0, k, FL_MOD[0], carry);
  r1 = mac(r1, k, FL_MOD[1], carry);
  r2 = mac(r2, k, FL_MOD[2], carry);
  r3 = mac(r3, k, FL_MOD[3], carry);
  { u128 t = (u128)r4 + carry; r4 = (uint64_t)t; carry2 = (uint64_t)(t >> 64); }

  k = r1 * FL_INV; carry = 0;
  (void)mac(r1, k, FL_MOD[0], carry);
  r2 = mac(r2, k, FL_MOD[1], carry);
  r3 = mac(r3, k, FL_MOD[2], carry);
  r4 = mac(r4, k, FL_MOD[3], carry);
  { u128 t = (u128)r5 + carry2 + carry; r5 = (uint64_t)t; carry2 = (uint64_t)(t >> 64); }

  k = r2 * FL_INV; carry = 0;
  (void)mac(r2, k, FL_MOD[0], carry);
  r3 = mac(r3, k, FL_MOD[1], carry);
  r4 = mac(r4, k, FL_MOD[2], carry);
  r5 = mac(r5, k, FL_MOD[3], carry);
  { u128 t = (u128)r6 + carry2 + carry; r6 = (uint64_t)t; carry2 = (uint64_t)(t >> 64); }

  k = r3 * FL_INV; carry = 0;
  (void)mac(r3, k, FL_MOD[0], carry);
  r4 = mac(r4, k, FL_MOD[1], carry);
  r5 = mac(r5, k, FL_MOD[2], carry);
  r6 = mac(r6, k, FL_MOD[3], carry);
  { u128 t = (u128)r7 + carry2 + carry; r7 = (uint64_t)t; }

  Fl r = {{r4, r5, r6, r7}};
  Fl m = {{FL_MOD[0], FL_MOD[1], FL_MOD[2], FL_MOD[3]}};
  return fl_sub(r, m);
}

// ristretto255.rs:702-726 (schoolbook 4x4 then reduce)
static inline Fl fl_mul(const Fl &a, const Fl &b) {
  uint64_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; j++) r[i + j] = mac(r[i + j], a.v[i], b.v[j], carry);
    r[i + 4] = carry;
  }
  return fl_mont_reduce(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
}
static inline Fl fl_sqr(const Fl &a) { return fl_mul(a, a); }

static inline Fl fl_from_u64(uint64_t x) { return fl_mul(Fl{{x, 0, 0, 0}}, FL_R2); }  // :213-217
static inline Fl fl_from_raw(const uint64_t x[4]) { return fl_mul(Fl{{x[0], x[1], x[2], x[3]}}, FL_R2); }

// :398-424; returns false if the 32 bytes are not canonical (>= l)
static inline bool fl_from_bytes(const uint8_t b[32], Fl *out) {
  Fl t;
  memcpy(t.v, b, 32);
  uint64_t br = 0;
  (void)sbb(t.v[0], FL_MOD[0], br);
  (void)sbb(t.v[1], FL_MOD[1], br);
  (void)sbb(t.v[2], FL_MOD[2], br);
  (void)sbb(t.v[3], FL_MOD[3], br);
  *out = fl_mul(t, FL_R2);
  return br == 1;
}
// :426-440
static inline void fl_to_bytes(const Fl &a, uint8_t out[32]) {
  Fl t = fl_mont_reduce(a.v[0], a.v[1], a.v[2], a.v[3], 0, 0, 0, 0);
  memcpy(out, t.v, 32);
}
// :442-473
static inline Fl fl_from_bytes_wide(const uint8_t b[64]) {
  Fl d0, d1;
  memcpy(d0.v, b, 32);
  memcpy(d1.v, b + 32, 32);
  return fl_add(fl_mul(d0, FL_R2), fl_mul(d1, FL_R3));
}

static inline Fl fl_pow(const Fl &a, const uint64_t e[4]) {
  Fl res = fl_one();
  for (int i = 3; i >= 0; i--)
    for (int j = 63; j >= 0; j--) {
      res = fl_sqr(res);
      if ((e[i] >> j) & 1) res = fl_mul(res, a);
    }
  return res;
}
// :548-602 computes a^(l-2) by an addition chain; the value is the same.
static inline Fl fl_invert(const Fl &a) {
  uint64_t e[4] = {FL_MOD[0] - 2, FL_MOD[1], FL_MOD[2], FL_MOD[3]};
  return fl_pow(a, e);
}
// :604-651 Montgomery's trick; returns the inverse of the product, inverts in place. Inputs must be non-zero.
static inline Fl fl_batch_invert(std::vector<Fl> &xs) {
  size_t n = xs.size();
  std::vector<Fl> scratch(n);
  Fl acc = fl_one();
  for (size_t i = 0; i < n; i++) { scratch[i] = acc; acc = fl_mul(acc, xs[i]); }
  acc = fl_invert(acc);
  Fl ret = acc;
  for (size_t i = n; i-- > 0;) {
    Fl tmp = fl_mul(acc, xs[i]);
    xs[i] = fl_mul(acc, scratch[i]);
    acc = tmp;
  }
  return ret;
}
static inline bool fl_is_zero(const Fl &a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }

}  // namespace orc
