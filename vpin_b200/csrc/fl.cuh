// F_l scalar field of ristretto255 for the B200 prover, l = 2^252 + 27742317777372353535851937790883648493.
// 8 x u32 limbs, Montgomery form with R = 2^256 — byte-identical in memory to the reference's 4 x u64 Montgomery
// `Scalar` (Spartan/src/scalar/ristretto255.rs:199-200), so device tables serialise straight into bincode proofs.
// Every operation returns the canonical representative in [0, l) (reference :654-698, :730-758).
// Same source compiles for host (orchestration) and device (kernels).
#pragma once
#include <cstdint>
#include <cstring>

#include "limbs.cuh"

namespace vpin {

struct alignas(16) fl_t { uint32_t v[8]; };

#define VPIN_FL_P0 0x5cf5d3edu
#define VPIN_FL_P1 0x5812631au
#define VPIN_FL_P2 0xa2f79cd6u
#define VPIN_FL_P3 0x14def9deu
#define VPIN_FL_P7 0x10000000u
#define VPIN_FL_INV32 0x12547e1bu  // -(l^-1) mod 2^32  (low half of reference INV, :305)

VPIN_HD fl_t fl_zero() { fl_t r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
// R = 2^256 mod l  (reference :308)
VPIN_HD fl_t fl_one() {
  fl_t r;
  r.v[0] = 0x8d98951du; r.v[1] = 0xd6ec3174u; r.v[2] = 0x737dcf70u; r.v[3] = 0xc6ef5bf4u;
  r.v[4] = 0xfffffffeu; r.v[5] = 0xffffffffu; r.v[6] = 0xffffffffu; r.v[7] = 0x0fffffffu;
  return r;
}
// R^2 mod l (reference :316)
VPIN_HD fl_t fl_r2() {
  fl_t r;
  r.v[0] = 0x449c0f01u; r.v[1] = 0xa40611e3u; r.v[2] = 0x68859347u; r.v[3] = 0xd00e1ba7u;
  r.v[4] = 0x17f5be65u; r.v[5] = 0xceec73d2u; r.v[6] = 0x7c309a3du; r.v[7] = 0x0399411bu;
  return r;
}
// R^3 mod l (reference :324)
VPIN_HD fl_t fl_r3() {
  fl_t r;
  r.v[0] = 0x7b83a2dbu; r.v[1] = 0x2a9e4968u; r.v[2] = 0xaef7f3ecu; r.v[3] = 0x278324e6u;
  r.v[4] = 0x04ec5b65u; r.v[5] = 0x8065dc6cu; r.v[6] = 0x3599cec7u; r.v[7] = 0x0e530b77u;
  return r;
}
VPIN_HD uint32_t fl_modulus_limb(int i) {
  switch (i) {
    case 0: return VPIN_FL_P0; case 1: return VPIN_FL_P1; case 2: return VPIN_FL_P2; case 3: return VPIN_FL_P3;
    case 7: return VPIN_FL_P7; default: return 0u;
  }
}
VPIN_HD bool fl_is_zero(const fl_t &a) {
  uint32_t o = 0;
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}
VPIN_HD bool fl_eq(const fl_t &a, const fl_t &b) {
  uint32_t o = 0;
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

// r = a - l if a >= l else a   (a < 2l)
VPIN_HD fl_t fl_cond_sub(const fl_t &a) {
  fl_t d;
#if defined(__CUDA_ARCH__)
  uint32_t br;
  asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;"
      "subc.cc.u32 %4, %13, 0; subc.cc.u32 %5, %14, 0; subc.cc.u32 %6, %15, 0; subc.cc.u32 %7, %16, %21; subc.u32 %8, 0, 0;"
      : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7]), "=r"(br)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "n"(VPIN_FL_P0), "n"(VPIN_FL_P1), "n"(VPIN_FL_P2), "n"(VPIN_FL_P3), "n"(VPIN_FL_P7));
  bool keep = br != 0;  // borrow -> a < l
#else
  int64_t br = 0;
  for (int i = 0; i < 8; i++) {
    int64_t t = (int64_t)a.v[i] - (int64_t)fl_modulus_limb(i) + br;
    d.v[i] = (uint32_t)t;
    br = t >> 32;
  }
  bool keep = br != 0;
#endif
  fl_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = keep ? a.v[i] : d.v[i];
  return r;
}
VPIN_HD fl_t fl_add(const fl_t &a, const fl_t &b) {
  fl_t s;
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, %23;"
      : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    s.v[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return fl_cond_sub(s);  // a + b < 2l < 2^254, no carry out
}
VPIN_HD fl_t fl_sub(const fl_t &a, const fl_t &b) {
  fl_t d;
#if defined(__CUDA_ARCH__)
  uint32_t mask;
  asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;"
      "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24; subc.u32 %8, 0, 0;"
      : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7]), "=r"(mask)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // mask = 0xffffffff on borrow: add l back
  asm("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11;"
      "addc.cc.u32 %4, %4, 0; addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, 0; addc.u32 %7, %7, %12;"
      : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]), "+r"(d.v[6]), "+r"(d.v[7])
      : "r"(mask & VPIN_FL_P0), "r"(mask & VPIN_FL_P1), "r"(mask & VPIN_FL_P2), "r"(mask & VPIN_FL_P3), "r"(mask & VPIN_FL_P7));
#else
  int64_t br = 0;
  for (int i = 0; i < 8; i++) {
    int64_t t = (int64_t)a.v[i] - (int64_t)b.v[i] + br;
    d.v[i] = (uint32_t)t;
    br = t >> 32;
  }
  uint32_t mask = br ? 0xffffffffu : 0u;
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)d.v[i] + (fl_modulus_limb(i) & mask);
    d.v[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return d;
}
VPIN_HD fl_t fl_neg(const fl_t &a) { return fl_sub(fl_zero(), a); }
VPIN_HD fl_t fl_dbl(const fl_t &a) { return fl_add(a, a); }

// Montgomery product a*b/R mod l: interleaved product / reduction rows of limbs.cuh (the modulus has limbs 4..6 == 0 and
// limb 7 == 2^28, so a reduction row costs 5 multiplies), then one conditional subtraction.
#if !defined(__CUDA_ARCH__)
// host: 4 x 64-bit CIOS Montgomery multiplication with 128-bit products (the transcript-side arithmetic of the prover)
inline fl_t fl_mul_host64(const fl_t &a, const fl_t &b) {
  typedef unsigned __int128 u128;
  static const uint64_t P[4] = {0x5812631a5cf5d3edull, 0x14def9dea2f79cd6ull, 0ull, 0x1000000000000000ull};
  static const uint64_t INV = 0xd2b51da312547e1bull;  // reference :305
  uint64_t x[4], y[4], t[6] = {0, 0, 0, 0, 0, 0};
  memcpy(x, a.v, 32);
  memcpy(y, b.v, 32);
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)x[j] * y[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * INV;
    c = (u128)m * P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  // t < 2l: one conditional subtraction
  uint64_t d[4];
  u128 br = 0;
  for (int j = 0; j < 4; j++) { u128 s = (u128)t[j] - P[j] - (uint64_t)br; d[j] = (uint64_t)s; br = (s >> 64) & 1; }
  fl_t r;
  memcpy(r.v, br ? t : d, 32);
  return r;
}
#endif
VPIN_HD fl_t fl_mul(const fl_t &a, const fl_t &b) {
#if defined(__CUDA_ARCH__)
  fl_t r;
  limb::mont_mul_l(r.v, a.v, b.v);
  return fl_cond_sub(r);
#else
  return fl_mul_host64(a, b);
#endif
}
VPIN_HD fl_t fl_sqr(const fl_t &a) { return fl_mul(a, a); }

// canonical (non-Montgomery) limbs of a: a * 1 / R
VPIN_HD fl_t fl_from_mont(const fl_t &a) {
#if defined(__CUDA_ARCH__)
  fl_t r;
  limb::mont_redc_l(r.v, a.v);  // the reduction rows of a multiplication by one, without its product rows
  return fl_cond_sub(r);
#else
  fl_t one;
  for (int i = 0; i < 8; i++) one.v[i] = 0;
  one.v[0] = 1;
  return fl_mul(a, one);
#endif
}
VPIN_HD fl_t fl_to_mont(const fl_t &a) { return fl_mul(a, fl_r2()); }
VPIN_HD fl_t fl_from_u64(uint64_t x) {
  fl_t t = fl_zero();
  t.v[0] = (uint32_t)x;
  t.v[1] = (uint32_t)(x >> 32);
  return fl_to_mont(t);
}

// ---- host-side helpers (byte formats of the reference API) ----
// reference from_bytes :398-424: rejects values >= l
inline bool fl_from_bytes(const uint8_t b[32], fl_t *out) {
  fl_t t;
  memcpy(t.v, b, 32);
  int64_t br = 0;
  for (int i = 0; i < 8; i++) {
    int64_t d = (int64_t)t.v[i] - (int64_t)fl_modulus_limb(i) + br;
    br = d >> 32;
  }
  *out = fl_to_mont(t);
  return br != 0;
}
inline void fl_to_bytes(const fl_t &a, uint8_t out[32]) {  // reference :426-440
  fl_t t = fl_from_mont(a);
  memcpy(out, t.v, 32);
}
inline fl_t fl_from_bytes_wide(const uint8_t b[64]) {  // reference :442-473
  fl_t d0, d1;
  memcpy(d0.v, b, 32);
  memcpy(d1.v, b + 32, 32);
  return fl_add(fl_mul(d0, fl_r2()), fl_mul(d1, fl_r3()));
}
VPIN_HD fl_t fl_pow_lm2(const fl_t &a) {  // a^(l-2): inversion (reference :548-602)
  // l - 2 limbs
  const uint32_t e[8] = {VPIN_FL_P0 - 2u, VPIN_FL_P1, VPIN_FL_P2, VPIN_FL_P3, 0u, 0u, 0u, VPIN_FL_P7};
  fl_t r = fl_one();
  for (int i = 7; i >= 0; i--)
    for (int j = 31; j >= 0; j--) {
      r = fl_sqr(r);
      if ((e[i] >> j) & 1) r = fl_mul(r, a);
    }
  return r;
}
VPIN_HD fl_t fl_invert(const fl_t &a) { return fl_pow_lm2(a); }

inline fl_t operator+(const fl_t &a, const fl_t &b) { return fl_add(a, b); }
inline fl_t operator-(const fl_t &a, const fl_t &b) { return fl_sub(a, b); }
inline fl_t operator*(const fl_t &a, const fl_t &b) { return fl_mul(a, b); }

}  // namespace vpin
