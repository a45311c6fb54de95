// F_p (p = 2^255 - 19) in NINE 29-bit limbs with 64-bit column accumulators: the carry-free multiplier of the MSM hot loop.
//
// Why a second representation next to ed.cuh's 8 x u32: in radix 2^32 every partial product has to travel through the carry
// flag (IMAD.WIDE.U32.X), and the carry-in form of IMAD.WIDE issues at HALF rate on sm_100a (scripts/ubench/imad_rates.cu):
// 2/3 of the multiplies of an 8 x 8 product are of that form. In radix 2^29 a column of the schoolbook product holds at most
// nine 58-bit products (< 2^62), so every partial product is a plain `mad.wide.s32` into a 64-bit column register - full rate,
// no flag, seventeen independent dependency chains per multiplication - and the carries are resolved once per
// multiplication by shifts and adds on the ALU pipe, which the chained form leaves idle.
//
// Representation: value = sum v[k] 2^(29 k), k = 0..8.
// Values are only reduced to 261 bits (2^261 = 2^6 2^255 == 19 * 64 = 1216 mod p folds columns 9..16 onto 0..8).
//   "normal"   : what f9_mul returns: every limb >= 0, v[1..8] < 2^29, v[0] < 2^29 + 2^24.
//   operands   : sums / differences of at most two normal values.
// Column bounds (N = 2^29; the 2^24 slack sits on limb 0 only and changes nothing below):
//   signed   : operands a in (-N, N), b in (-2N, 2N): |column| <= 9 * 2 N^2 = 2^62.17; fold (< 2^41), bias (2^41) and carries
//              (< 2^34) keep it below 2^63.
//   unsigned : both operands in [0, 2N) (the G * H product of a mixed addition): column <= 9 * 4 N^2 = 2^63.17 < 2^64.
// Host and device share this source (the host build is what tests/cpp/test_fp29.cpp checks against big integers).
#pragma once
#include "ed.cuh"

namespace vpin {

struct f9 { int32_t v[9]; };

static const int32_t kF9Mask = (1 << 29) - 1;

VPIN_HD int64_t f9_mulw_s(int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
  int64_t r;
  asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
  return r;
#else
  return (int64_t)a * b;
#endif
}
VPIN_HD void f9_madw_s(int64_t &acc, int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
  // written as a carry-flag pair on purpose: ptxas fuses the pair into ONE IMAD.WIDE with a 64-bit addend and leaves it
  // alone, whereas chains of plain `mad.wide` get re-associated into independent products + 3-input 64-bit adds (IADD3 /
  // IADD3.X pairs on the ALU pipe: +60 % instructions, measured)
  uint32_t lo = (uint32_t)acc, hi = (uint32_t)((uint64_t)acc >> 32);
  asm("mad.lo.cc.s32 %0, %2, %3, %0; madc.hi.s32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
  acc = (int64_t)(((uint64_t)hi << 32) | lo);
#else
  acc += (int64_t)a * b;
#endif
}
VPIN_HD uint64_t f9_mulw_u(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint64_t r;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
  return r;
#else
  return (uint64_t)a * b;
#endif
}
VPIN_HD void f9_madw_u(uint64_t &acc, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
  asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
  acc = ((uint64_t)hi << 32) | lo;
#else
  acc += (uint64_t)a * b;
#endif
}

// plain mad.wide: ptxas re-associates chains of these into independent IMAD.WIDE products (no addend) summed by three-input
// 64-bit adds (IADD3 / IADD3.X with two carry flags) on the ALU pipe
VPIN_HD void f9_madw_s_alu(int64_t &acc, int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
  acc += (int64_t)a * b;
#endif
}
VPIN_HD void f9_madw_u_alu(uint64_t &acc, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
  acc += (uint64_t)a * b;
#endif
}

// 8 x u32 (value < 2^256) -> nine 29-bit limbs (the top one holds bits 232..255)
VPIN_HD f9 f9_unpack(const uint32_t w[8]) {
  f9 r;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int o = 29 * k, j = o >> 5, sh = o & 31;
    uint32_t lo = w[j], hi = (sh + 29 > 32) ? w[j + 1] : 0u;
    uint32_t x = sh == 0 ? lo : ((lo >> sh) | (sh + 29 > 32 ? (hi << (32 - sh)) : 0u));
    r.v[k] = (int32_t)(x & (uint32_t)kF9Mask);
  }
  r.v[8] = (int32_t)(w[7] >> 8);
  return r;
}
VPIN_HD f9 f9_from_fp(const fp_t &a) { return f9_unpack(a.v); }
// any value with non-negative limbs < 2^31 -> 8 x u32, lazily reduced below 2^256 like ed.cuh's fp_t
VPIN_HD fp_t f9_to_fp(const f9 &a) {
  // exact 29-bit limbs, fold what lies above bit 255 (2^255 == 19), exact limbs again, pack
  uint32_t l[9];
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    c += (uint64_t)(uint32_t)a.v[k];
    l[k] = k < 8 ? (uint32_t)(c & (uint32_t)kF9Mask) : (uint32_t)c;
    c >>= 29;
  }
  c = 19ull * (l[8] >> 23);
  l[8] &= (1u << 23) - 1u;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    c += l[k];
    l[k] = k < 8 ? (uint32_t)(c & (uint32_t)kF9Mask) : (uint32_t)c;
    c >>= 29;
  }
  fp_t r;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    // word j holds bits 32 j .. 32 j + 31
    const int lo_k = (32 * j) / 29, lo_sh = 32 * j - 29 * lo_k;
    uint64_t x = (uint64_t)l[lo_k] >> lo_sh;
    int have = 29 - lo_sh;
    if (lo_k + 1 < 9) { x |= (uint64_t)l[lo_k + 1] << have; have += 29; }
    if (have < 32 && lo_k + 2 < 9) x |= (uint64_t)l[lo_k + 2] << have;
    r.v[j] = (uint32_t)x;
  }
  return r;
}
VPIN_HD f9 f9_zero() { f9 r; for (int k = 0; k < 9; k++) r.v[k] = 0; return r; }
VPIN_HD f9 f9_one() { f9 r = f9_zero(); r.v[0] = 1; return r; }
VPIN_HD f9 f9_add(const f9 &a, const f9 &b) { f9 r; for (int k = 0; k < 9; k++) r.v[k] = a.v[k] + b.v[k]; return r; }
VPIN_HD f9 f9_sub(const f9 &a, const f9 &b) { f9 r; for (int k = 0; k < 9; k++) r.v[k] = a.v[k] - b.v[k]; return r; }
VPIN_HD f9 f9_cneg(const f9 &a, bool neg) { f9 r; for (int k = 0; k < 9; k++) r.v[k] = neg ? -a.v[k] : a.v[k]; return r; }

// a * b mod p -> normal form. kUnsigned: both operands have non-negative limbs (columns may then reach 2^63..2^64).
// Order of the carries: the top of column 8 first joins the high half (columns 9..16 -> h[0..8], exact 29-bit limbs), the high
// half is folded onto columns 0..8 with 2^261 == 1216, then ONE pass over the low columns; what leaves column 8 at the end is
// small (< 2^13) and goes to limb 0 times 1216.
// kPolicy: where the 64 accumulations of a product happen. An IMAD.WIDE with a 64-bit addend occupies the multiply pipe for
// twice as long as one without (measured, scripts/ubench/imad_rates2.cu); the alternative is a plain product and a 64-bit add
// on the ALU pipe (IADD3 + IADD3.X take three inputs: one pair absorbs two products).
//   0: every accumulation in the multiplier (carry-flag pairs)   1: every accumulation on the ALU pipe
VPIN_HD constexpr bool f9_on_alu(int policy, int t) {
  return policy == 1;
}
template <bool kUnsigned, int kPolicy = 0>
VPIN_HD f9 f9_mul(const f9 &a, const f9 &b) {
  f9 r;
  if (kUnsigned) {
    uint64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) {
      int t = 0;
#pragma unroll
      for (int i = 0; i < 9; i++) {
        const int j = k - i;
        if (j < 0 || j > 8) continue;
        if (t == 0) c[k] = f9_mulw_u((uint32_t)a.v[i], (uint32_t)b.v[j]);
        else if (f9_on_alu(kPolicy, t)) f9_madw_u_alu(c[k], (uint32_t)a.v[i], (uint32_t)b.v[j]);
        else f9_madw_u(c[k], (uint32_t)a.v[i], (uint32_t)b.v[j]);
        t++;
      }
    }
    uint32_t h[9];
    c[9] += c[8] >> 29;
    c[8] &= (uint64_t)kF9Mask;
#pragma unroll
    for (int k = 9; k < 16; k++) { c[k + 1] += c[k] >> 29; h[k - 9] = (uint32_t)c[k] & (uint32_t)kF9Mask; }
    h[7] = (uint32_t)c[16] & (uint32_t)kF9Mask;
    h[8] = (uint32_t)(c[16] >> 29);
#pragma unroll
    for (int k = 0; k < 9; k++) f9_madw_u(c[k], h[k], 1216u);
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k + 1] += c[k] >> 29; r.v[k] = (int32_t)((uint32_t)c[k] & (uint32_t)kF9Mask); }
    r.v[8] = (int32_t)((uint32_t)c[8] & (uint32_t)kF9Mask);
    r.v[0] += (int32_t)(1216u * (uint32_t)(c[8] >> 29));
  } else {
    int64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) {
      int t = 0;
#pragma unroll
      for (int i = 0; i < 9; i++) {
        const int j = k - i;
        if (j < 0 || j > 8) continue;
        if (t == 0) c[k] = f9_mulw_s(a.v[i], b.v[j]);
        else if (f9_on_alu(kPolicy, t)) f9_madw_s_alu(c[k], a.v[i], b.v[j]);
        else f9_madw_s(c[k], a.v[i], b.v[j]);
        t++;
      }
    }
    int32_t h[9];
    c[9] += c[8] >> 29;
    c[8] &= (int64_t)kF9Mask;
#pragma unroll
    for (int k = 9; k < 16; k++) { c[k + 1] += c[k] >> 29; h[k - 9] = (int32_t)c[k] & kF9Mask; }
    h[7] = (int32_t)c[16] & kF9Mask;
    h[8] = (int32_t)(c[16] >> 29);
#pragma unroll
    for (int k = 0; k < 9; k++) f9_madw_s(c[k], h[k], 1216);
    // positivity bias 2^18 p = 2^41 2^232 - 19 2^18: after the fold |column 8| < 2^40.5, so + 2^41 there (and - 19 2^18 at the
    // bottom) makes the carry out of column 8 - and with it every limb of the result - non-negative
    c[0] -= (int64_t)19 << 18;
    c[8] += (int64_t)1 << 41;
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k + 1] += c[k] >> 29; r.v[k] = (int32_t)c[k] & kF9Mask; }
    r.v[8] = (int32_t)c[8] & kF9Mask;
    r.v[0] += 1216 * (int32_t)(c[8] >> 29);
  }
  return r;
}

}  // namespace vpin
