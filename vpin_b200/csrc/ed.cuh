// F_p (p = 2^255 - 19) and the ristretto255 group for the B200 prover; host + device.
// Replaces what the reference reaches through curve25519-dalek 3.2.0 (Spartan/src/group.rs:6-8, :103-121;
// Spartan/src/commitments.rs:30 from_uniform_bytes; every .compress()/.decompress()).
// Field elements are 8 x u32 limbs kept lazily reduced in [0, 2^256) (2^256 == 38 mod p); they are made canonical
// only when bytes are produced. Points are extended twisted Edwards (a = -1). Encoding is RFC 9496, which is
// what dalek's CompressedRistretto is, so equal group elements give equal bytes whatever formulas were used.
#pragma once
#include "fl.cuh"

namespace vpin {

struct alignas(16) fp_t { uint32_t v[8]; };

#define VPIN_FP_D {{0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu}}
#define VPIN_FP_D2 {{0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu}}
#define VPIN_FP_SQRT_M1 {{0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u}}
#define VPIN_FP_SQRT_AD_MINUS_ONE {{0x497b2e1bu, 0x7e97f6a0u, 0x1b7854bdu, 0xaf9d8e0cu, 0x31f5d1fdu, 0x0f3cfcc9u, 0x2b8348acu, 0x376931bfu}}
#define VPIN_FP_INVSQRT_A_MINUS_D {{0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u}}
#define VPIN_FP_ONE_MINUS_D_SQ {{0x945fc176u, 0xe27c09c1u, 0xcd5e350fu, 0x2c81a138u, 0xbe70dfe4u, 0x9994abddu, 0xb2b3e0d7u, 0x029072a8u}}
#define VPIN_FP_D_MINUS_ONE_SQ {{0x44ed4d20u, 0x31ad5aaau, 0xb01e1999u, 0xd29e4a2cu, 0x529b4eebu, 0x4cdcd32fu, 0xf66c2241u, 0x5968b37au}}

VPIN_HD fp_t fp_d() { const fp_t c = VPIN_FP_D; return c; }
VPIN_HD fp_t fp_d2() { const fp_t c = VPIN_FP_D2; return c; }
VPIN_HD fp_t fp_sqrt_m1() { const fp_t c = VPIN_FP_SQRT_M1; return c; }
VPIN_HD fp_t fp_sqrt_ad_minus_one() { const fp_t c = VPIN_FP_SQRT_AD_MINUS_ONE; return c; }
VPIN_HD fp_t fp_invsqrt_a_minus_d() { const fp_t c = VPIN_FP_INVSQRT_A_MINUS_D; return c; }
VPIN_HD fp_t fp_one_minus_d_sq() { const fp_t c = VPIN_FP_ONE_MINUS_D_SQ; return c; }
VPIN_HD fp_t fp_d_minus_one_sq() { const fp_t c = VPIN_FP_D_MINUS_ONE_SQ; return c; }

VPIN_HD fp_t fp_zero() { fp_t r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
VPIN_HD fp_t fp_one() { fp_t r = fp_zero(); r.v[0] = 1; return r; }

VPIN_HD fp_t fp_add(const fp_t &a, const fp_t &b) {
  fp_t r;
#if defined(__CUDA_ARCH__)
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24; addc.u32 %8, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(c)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // fold the carry twice (2^256 == 38)
  asm("mad.lo.cc.u32 %0, %8, 38, %0; addc.cc.u32 %1, %1, 0; addc.cc.u32 %2, %2, 0; addc.cc.u32 %3, %3, 0;"
      "addc.cc.u32 %4, %4, 0; addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0; addc.u32 %8, 0, 0;"
      "mad.lo.u32 %0, %8, 38, %0;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "+r"(c));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  c *= 38;
  for (int i = 0; i < 8; i++) { c += r.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  r.v[0] += 38u * (uint32_t)c;
#endif
  return r;
}
VPIN_HD fp_t fp_sub(const fp_t &a, const fp_t &b) {
  fp_t r;
#if defined(__CUDA_ARCH__)
  uint32_t br;
  asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;"
      "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24; subc.u32 %8, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(br)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // br = 0xffffffff when the difference wrapped by 2^256: subtract 38, possibly twice
  br &= 38u;
  asm("sub.cc.u32 %0, %0, %8; subc.cc.u32 %1, %1, 0; subc.cc.u32 %2, %2, 0; subc.cc.u32 %3, %3, 0;"
      "subc.cc.u32 %4, %4, 0; subc.cc.u32 %5, %5, 0; subc.cc.u32 %6, %6, 0; subc.cc.u32 %7, %7, 0; subc.u32 %8, 0, 0;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "+r"(br));
  r.v[0] -= br & 38u;
#else
  int64_t br = 0;
  for (int i = 0; i < 8; i++) { int64_t t = (int64_t)a.v[i] - (int64_t)b.v[i] + br; r.v[i] = (uint32_t)t; br = t >> 32; }
  // wrapped by 2^256 -> subtract 38, possibly twice
  int64_t k = br ? 38 : 0;
  br = 0;
  for (int i = 0; i < 8; i++) { int64_t t = (int64_t)r.v[i] - (i == 0 ? k : 0) + br; r.v[i] = (uint32_t)t; br = t >> 32; }
  r.v[0] -= br ? 38u : 0u;
#endif
  return r;
}
VPIN_HD fp_t fp_neg(const fp_t &a) { return fp_sub(fp_zero(), a); }
// a / 2: (a + (a odd ? p : 0)) >> 1, again lazily reduced below 2^256
VPIN_HD fp_t fp_half(const fp_t &a) {
  const uint32_t odd = 0u - (a.v[0] & 1u);
  uint32_t t[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + ((i == 0 ? 0xffffffedu : (i == 7 ? 0x7fffffffu : 0xffffffffu)) & odd);
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  t[8] = (uint32_t)c;
  fp_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
  return r;
}

// t (16 limbs) -> t mod 2^256-38, lazily reduced into [0, 2^256)
VPIN_HD fp_t fp_reduce_wide(const uint32_t t[16]) {
  fp_t r;
#if defined(__CUDA_ARCH__)
  uint32_t lo[8], tmp[8], w8 = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) lo[i] = t[i];
  limb::mad_row(lo, t + 8, 38u, w8);   // 38 * (t8, t10, t12, t14) at limbs 0, 2, 4, 6
  limb::mul_row(tmp, t + 9, 38u);      // 38 * (t9, t11, t13, t15) at limbs 1, 3, 5, 7
  asm("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11;"
      "addc.cc.u32 %4, %4, %12; addc.cc.u32 %5, %5, %13; addc.cc.u32 %6, %6, %14; addc.u32 %7, %7, %15;"
      : "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]), "+r"(w8)
      : "r"(tmp[0]), "r"(tmp[1]), "r"(tmp[2]), "r"(tmp[3]), "r"(tmp[4]), "r"(tmp[5]), "r"(tmp[6]), "r"(tmp[7]));
  // w8 <= 39: fold it, and the (rare) carry of that fold
  asm("mad.lo.cc.u32 %0, %8, 38, %0; addc.cc.u32 %1, %1, 0; addc.cc.u32 %2, %2, 0; addc.cc.u32 %3, %3, 0;"
      "addc.cc.u32 %4, %4, 0; addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0; addc.u32 %8, 0, 0;"
      "mad.lo.u32 %0, %8, 38, %0;"
      : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]), "+r"(w8));
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = lo[i];
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)t[8 + i] * 38u + t[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  c *= 38;
  for (int i = 0; i < 8; i++) { c += r.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  r.v[0] += 38u * (uint32_t)c;
#endif
  return r;
}
VPIN_HD fp_t fp_mul(const fp_t &a, const fp_t &b) {
  uint32_t t[16];
  limb::mul_8x8(t, a.v, b.v);
  return fp_reduce_wide(t);
}
// a * a with the dedicated squaring of limbs.cuh (36 + 8 IMAD.WIDE instead of 64 + 8); same value as fp_mul(a, a)
VPIN_HD fp_t fp_sqr(const fp_t &a) {
  uint32_t t[16];
  limb::sqr_8x8(t, a.v);
  return fp_reduce_wide(t);
}
// fp_mul with the accumulation of chosen product rows moved to the ALU pipe (limbs.cuh mul_8x8_p); same value
template <uint32_t kAluRows>
VPIN_HD fp_t fp_mul_p(const fp_t &a, const fp_t &b) {
  uint32_t t[16];
  limb::mul_8x8_p<kAluRows>(t, a.v, b.v);
  return fp_reduce_wide(t);
}

// canonical bytes (value in [0, p))
VPIN_HD void fp_canon(const fp_t &a, uint32_t out[8]) {
  // x = (a mod 2^255) + 19 * (a >> 255)  < 2^255 + 19
  uint32_t x[8];
  uint64_t c = 19ull * (a.v[7] >> 31);
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (i == 7 ? (a.v[7] & 0x7fffffffu) : a.v[i]); x[i] = (uint32_t)c; c >>= 32; }
  // y = x + 19; if y >= 2^255 then x >= p: result y - 2^255
  uint32_t y[8];
  c = 19;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += x[i]; y[i] = (uint32_t)c; c >>= 32; }
  bool ge = (y[7] >> 31) != 0;
  y[7] &= 0x7fffffffu;
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = ge ? y[i] : x[i];
}
VPIN_HD bool fp_is_neg(const fp_t &a) { uint32_t c[8]; fp_canon(a, c); return c[0] & 1; }
VPIN_HD bool fp_is_zero(const fp_t &a) {
  uint32_t c[8];
  fp_canon(a, c);
  uint32_t o = 0;
  for (int i = 0; i < 8; i++) o |= c[i];
  return o == 0;
}
VPIN_HD bool fp_eq(const fp_t &a, const fp_t &b) { return fp_is_zero(fp_sub(a, b)); }
VPIN_HD fp_t fp_abs(const fp_t &a) { return fp_is_neg(a) ? fp_neg(a) : a; }
VPIN_HD fp_t fp_sqr_n(fp_t a, int n) { for (int i = 0; i < n; i++) a = fp_sqr(a); return a; }
// a^(2^252 - 3) = a^((p-5)/8), standard addition chain (11 multiplications)
VPIN_HD fp_t fp_pow22523(const fp_t &z) {
  fp_t t0 = fp_sqr(z);                       // 2
  fp_t t1 = fp_mul(z, fp_sqr_n(t0, 2));      // 9
  t0 = fp_mul(t0, t1);                       // 11
  t0 = fp_mul(t1, fp_sqr(t0));               // 31 = 2^5-1
  t1 = fp_mul(fp_sqr_n(t0, 5), t0);          // 2^10-1
  fp_t t2 = fp_mul(fp_sqr_n(t1, 10), t1);    // 2^20-1
  fp_t t3 = fp_mul(fp_sqr_n(t2, 20), t2);    // 2^40-1
  t2 = fp_mul(fp_sqr_n(t3, 10), t1);         // 2^50-1
  t3 = fp_mul(fp_sqr_n(t2, 50), t2);         // 2^100-1
  fp_t t4 = fp_mul(fp_sqr_n(t3, 100), t3);   // 2^200-1
  t3 = fp_mul(fp_sqr_n(t4, 50), t2);         // 2^250-1
  return fp_mul(fp_sqr_n(t3, 2), z);         // 2^252-3
}
// a^(p-2)
VPIN_HD fp_t fp_invert(const fp_t &z) {
  // z^(2^255-21) = (z^(2^252-3))^8 * z^3
  fp_t t = fp_sqr_n(fp_pow22523(z), 3);
  return fp_mul(t, fp_mul(fp_sqr(z), z));
}
VPIN_HD fp_t fp_from_bytes(const uint8_t b[32]) {  // ignores bit 255 (dalek FieldElement::from_bytes)
  fp_t r;
#pragma unroll
  for (int i = 0; i < 8; i++)
    r.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  r.v[7] &= 0x7fffffffu;
  return r;
}
VPIN_HD void fp_to_bytes(const fp_t &a, uint8_t out[32]) {
  uint32_t c[8];
  fp_canon(a, c);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)c[i]; out[4 * i + 1] = (uint8_t)(c[i] >> 8);
    out[4 * i + 2] = (uint8_t)(c[i] >> 16); out[4 * i + 3] = (uint8_t)(c[i] >> 24);
  }
}

// RFC 9496 4.2 SQRT_RATIO_M1
VPIN_HD bool fp_sqrt_ratio_m1(const fp_t &u, const fp_t &v, fp_t *out) {
  fp_t v3 = fp_mul(fp_sqr(v), v);
  fp_t v7 = fp_mul(fp_sqr(v3), v);
  fp_t r = fp_mul(fp_mul(u, v3), fp_pow22523(fp_mul(u, v7)));
  fp_t check = fp_mul(v, fp_sqr(r));
  fp_t neg_u = fp_neg(u);
  bool correct = fp_eq(check, u);
  bool flipped = fp_eq(check, neg_u);
  bool flipped_i = fp_eq(check, fp_mul(neg_u, fp_sqrt_m1()));
  if (flipped || flipped_i) r = fp_mul(r, fp_sqrt_m1());
  *out = fp_abs(r);
  return correct || flipped;
}

// ---- points ----
struct ge_t { fp_t X, Y, Z, T; };           // extended coordinates
struct niels_t { fp_t yp, ym, t2d; };       // affine "Niels": y+x, y-x, 2d*x*y (96 bytes, the MSM table entry)

VPIN_HD ge_t ge_identity() { ge_t r; r.X = fp_zero(); r.Y = fp_one(); r.Z = fp_one(); r.T = fp_zero(); return r; }
VPIN_HD ge_t ge_neg(const ge_t &p) { ge_t r; r.X = fp_neg(p.X); r.Y = p.Y; r.Z = p.Z; r.T = fp_neg(p.T); return r; }
// mixed addition, 7 field multiplications (Hisil-Wong-Carter-Dawson, a = -1, Z2 = 1)
VPIN_HD ge_t ge_madd(const ge_t &p, const niels_t &q) {
  fp_t a = fp_mul(fp_sub(p.Y, p.X), q.ym);
  fp_t b = fp_mul(fp_add(p.Y, p.X), q.yp);
  fp_t c = fp_mul(p.T, q.t2d);
  fp_t d = fp_add(p.Z, p.Z);
  fp_t e = fp_sub(b, a), f = fp_sub(d, c), g = fp_add(d, c), h = fp_add(b, a);
  ge_t r;
  r.X = fp_mul(e, f); r.Y = fp_mul(g, h); r.Z = fp_mul(f, g); r.T = fp_mul(e, h);
  return r;
}
VPIN_HD ge_t ge_msub(const ge_t &p, const niels_t &q) {
  niels_t n;
  n.yp = q.ym; n.ym = q.yp; n.t2d = fp_neg(q.t2d);
  return ge_madd(p, n);
}
// full addition, 9 multiplications
VPIN_HD ge_t ge_add(const ge_t &p, const ge_t &q) {
  fp_t a = fp_mul(fp_sub(p.Y, p.X), fp_sub(q.Y, q.X));
  fp_t b = fp_mul(fp_add(p.Y, p.X), fp_add(q.Y, q.X));
  fp_t c = fp_mul(fp_mul(p.T, q.T), fp_d2());
  fp_t d = fp_mul(p.Z, q.Z);
  d = fp_add(d, d);
  fp_t e = fp_sub(b, a), f = fp_sub(d, c), g = fp_add(d, c), h = fp_add(b, a);
  ge_t r;
  r.X = fp_mul(e, f); r.Y = fp_mul(g, h); r.Z = fp_mul(f, g); r.T = fp_mul(e, h);
  return r;
}
VPIN_HD ge_t ge_dbl(const ge_t &p) {
  fp_t a = fp_sqr(p.X), b = fp_sqr(p.Y), c = fp_sqr(p.Z);
  c = fp_add(c, c);
  fp_t d = fp_neg(a);
  fp_t xy = fp_add(p.X, p.Y);
  fp_t e = fp_sub(fp_sub(fp_sqr(xy), a), b);
  fp_t g = fp_add(d, b), f = fp_sub(g, c), h = fp_sub(d, b);
  ge_t r;
  r.X = fp_mul(e, f); r.Y = fp_mul(g, h); r.Z = fp_mul(f, g); r.T = fp_mul(e, h);
  return r;
}
// affine Niels form of p given zinv = 1/Z
VPIN_HD niels_t ge_to_niels(const ge_t &p, const fp_t &zinv) {
  fp_t x = fp_mul(p.X, zinv), y = fp_mul(p.Y, zinv);
  niels_t n;
  n.yp = fp_add(y, x); n.ym = fp_sub(y, x); n.t2d = fp_mul(fp_mul(x, y), fp_d2());
  return n;
}
// the same entry with every coordinate halved ((y+x)/2, (y-x)/2, d*x*y): what the MSM tables hold. With halved entries the
// mixed addition uses D = Z1 instead of 2 Z1 and returns (X3/4 : Y3/4 : Z3/4 : T3/4), the same projective point; the
// radix-2^29 multiplier of the hot loop (fp29.cuh) needs the smaller F = D - C, G = D + C this gives.
VPIN_HD niels_t niels_half(const niels_t &n) {
  niels_t h;
  h.yp = fp_half(n.yp); h.ym = fp_half(n.ym); h.t2d = fp_half(n.t2d);
  return h;
}
// RFC 9496 4.3.2 ENCODE  (dalek RistrettoPoint::compress)
VPIN_HD void ge_compress(const ge_t &p, uint8_t out[32]) {
  fp_t u1 = fp_mul(fp_add(p.Z, p.Y), fp_sub(p.Z, p.Y));
  fp_t u2 = fp_mul(p.X, p.Y);
  fp_t invsqrt;
  fp_sqrt_ratio_m1(fp_one(), fp_mul(u1, fp_sqr(u2)), &invsqrt);
  fp_t den1 = fp_mul(invsqrt, u1), den2 = fp_mul(invsqrt, u2);
  fp_t z_inv = fp_mul(fp_mul(den1, den2), p.T);
  fp_t ix0 = fp_mul(p.X, fp_sqrt_m1()), iy0 = fp_mul(p.Y, fp_sqrt_m1());
  fp_t enchanted = fp_mul(den1, fp_invsqrt_a_minus_d());
  bool rotate = fp_is_neg(fp_mul(p.T, z_inv));
  fp_t x = rotate ? iy0 : p.X, y = rotate ? ix0 : p.Y, den_inv = rotate ? enchanted : den2;
  if (fp_is_neg(fp_mul(x, z_inv))) y = fp_neg(y);
  fp_t s = fp_abs(fp_mul(den_inv, fp_sub(p.Z, y)));
  fp_to_bytes(s, out);
}
// RFC 9496 4.3.1 DECODE (dalek CompressedRistretto::decompress)
VPIN_HD bool ge_decompress(const uint8_t in[32], ge_t *out) {
  fp_t s = fp_from_bytes(in);
  uint8_t chk[32];
  fp_to_bytes(s, chk);
  bool canon = true;
  for (int i = 0; i < 32; i++) canon = canon && (chk[i] == in[i]);
  if (!canon || (in[0] & 1)) return false;
  fp_t one = fp_one();
  fp_t ss = fp_sqr(s);
  fp_t u1 = fp_sub(one, ss), u2 = fp_add(one, ss);
  fp_t u2_sqr = fp_sqr(u2);
  fp_t v = fp_sub(fp_neg(fp_mul(fp_d(), fp_sqr(u1))), u2_sqr);
  fp_t invsqrt;
  bool ok = fp_sqrt_ratio_m1(one, fp_mul(v, u2_sqr), &invsqrt);
  fp_t den_x = fp_mul(invsqrt, u2);
  fp_t den_y = fp_mul(fp_mul(invsqrt, den_x), v);
  fp_t x = fp_abs(fp_mul(fp_add(s, s), den_x));
  fp_t y = fp_mul(u1, den_y);
  fp_t t = fp_mul(x, y);
  if (!ok || fp_is_neg(t) || fp_is_zero(y)) return false;
  out->X = x; out->Y = y; out->Z = one; out->T = t;
  return true;
}
// RFC 9496 4.3.4 MAP
VPIN_HD ge_t ge_elligator(const fp_t &t) {
  fp_t one = fp_one();
  fp_t r = fp_mul(fp_sqrt_m1(), fp_sqr(t));
  fp_t u = fp_mul(fp_add(r, one), fp_one_minus_d_sq());
  fp_t v = fp_mul(fp_sub(fp_neg(one), fp_mul(r, fp_d())), fp_add(r, fp_d()));
  fp_t s;
  bool was_square = fp_sqrt_ratio_m1(u, v, &s);
  fp_t s_prime = fp_neg(fp_abs(fp_mul(s, t)));
  if (!was_square) s = s_prime;
  fp_t c = was_square ? fp_neg(one) : r;
  fp_t n = fp_sub(fp_mul(fp_mul(c, fp_sub(r, one)), fp_d_minus_one_sq()), v);
  fp_t w0 = fp_mul(fp_add(s, s), v);
  fp_t w1 = fp_mul(n, fp_sqrt_ad_minus_one());
  fp_t ss = fp_sqr(s);
  fp_t w2 = fp_sub(one, ss), w3 = fp_add(one, ss);
  ge_t g;
  g.X = fp_mul(w0, w3); g.Y = fp_mul(w2, w1); g.Z = fp_mul(w1, w3); g.T = fp_mul(w0, w2);
  return g;
}
// dalek RistrettoPoint::from_uniform_bytes (Spartan/src/commitments.rs:30)
VPIN_HD ge_t ge_from_uniform_bytes(const uint8_t b[64]) {
  return ge_add(ge_elligator(fp_from_bytes(b)), ge_elligator(fp_from_bytes(b + 32)));
}

}  // namespace vpin
