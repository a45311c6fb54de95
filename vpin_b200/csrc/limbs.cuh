// Multi-limb building blocks shared by F_p (ed.cuh) and F_l (fl.cuh): 256 x 256 -> 512-bit products in 32-bit limbs.
//
// Device path: every 32 x 32 -> 64 partial product is ONE IMAD.WIDE.U32 with carry-in/carry-out (ptxas fuses the
// mad.lo.cc / madc.hi.cc pairs below into IMAD.WIDE.U32[.X] Rd, P, Ra, Rb, Rc, P), using the even/odd column split:
// the products a[j] * b of one row with j even occupy disjoint 64-bit slots and are chained by the carry flag, the
// ones with j odd go to a second accumulator that is offset by one limb. 64 IMAD.WIDE per 8 x 8 product instead of
// the ~190 IMAD/IADD3 the compiler emits for the portable `c += (uint64_t)a * b + t` loop.
// Host path: the same functions in portable 64-bit arithmetic (used by the host-side sigma protocols and by the
// CPU-side unit tests of the index logic).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define VPIN_HD __host__ __device__ __forceinline__
#else
#define VPIN_HD inline
#endif

namespace vpin {
namespace limb {

// r[0..8) = (a[0], a[2], a[4], a[6]) * b as four 64-bit products (a is addressed with stride 2)
VPIN_HD void mul_row(uint32_t *r, const uint32_t *a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(r[0]), "=r"(r[1]) : "r"(a[0]), "r"(b));
  asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(r[2]), "=r"(r[3]) : "r"(a[2]), "r"(b));
  asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(r[4]), "=r"(r[5]) : "r"(a[4]), "r"(b));
  asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(r[6]), "=r"(r[7]) : "r"(a[6]), "r"(b));
#else
  for (int k = 0; k < 4; k++) {
    uint64_t p = (uint64_t)a[2 * k] * b;
    r[2 * k] = (uint32_t)p;
    r[2 * k + 1] = (uint32_t)(p >> 32);
  }
#endif
}
// r[0..8) += (a[0], a[2], a[4], a[6]) * b with one carry chain; cw += carry out
VPIN_HD void mad_row(uint32_t *r, const uint32_t *a, uint32_t b, uint32_t &cw) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;"
      "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;"
      "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;"
      "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7;"
      "addc.u32 %8, %8, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(cw)
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
#else
  uint64_t c = 0;
  for (int k = 0; k < 4; k++) {
    uint64_t p = (uint64_t)a[2 * k] * b;
    uint64_t lo = (uint64_t)r[2 * k] + (uint32_t)p + c;
    r[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)r[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    r[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  cw += (uint32_t)c;
#endif
}
// same without a carry word (the caller knows the chain cannot overflow)
VPIN_HD void mad_row_nc(uint32_t *r, const uint32_t *a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1;"
      "madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
      "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5;"
      "madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
#else
  uint32_t dummy = 0;
  mad_row(r, a, b, dummy);
#endif
}

// r[0..8) += (a[0], a[2], a[4], a[6]) * b like mad_row, but as four PLAIN products (IMAD.WIDE without addend) and one
// 9-word carry chain of IADD3.X. An IMAD.WIDE that also accumulates (64-bit addend, with or without carry-in) occupies the
// multiply pipe twice as long as a plain product (scripts/ubench/imad_rates2.cu), while the ALU pipe, which issues an IADD3
// every cycle, is otherwise idle: rows written this way trade 8 multiply-pipe cycles for 9 issue slots.
VPIN_HD void mad_row_alu(uint32_t *r, const uint32_t *a, uint32_t b, uint32_t &cw) {
#if defined(__CUDA_ARCH__)
  uint64_t p0, p1, p2, p3;  // (mul.wide, not mul.lo / mul.hi pairs: those ptxas fuses with the adds below into IMAD.WIDE.X again)
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a[0]), "r"(b));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(a[2]), "r"(b));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p2) : "r"(a[4]), "r"(b));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p3) : "r"(a[6]), "r"(b));
  uint32_t q[8] = {(uint32_t)p0, (uint32_t)(p0 >> 32), (uint32_t)p1, (uint32_t)(p1 >> 32),
                   (uint32_t)p2, (uint32_t)(p2 >> 32), (uint32_t)p3, (uint32_t)(p3 >> 32)};
  asm("add.cc.u32 %0, %0, %9; addc.cc.u32 %1, %1, %10; addc.cc.u32 %2, %2, %11; addc.cc.u32 %3, %3, %12;"
      "addc.cc.u32 %4, %4, %13; addc.cc.u32 %5, %5, %14; addc.cc.u32 %6, %6, %15; addc.cc.u32 %7, %7, %16; addc.u32 %8, %8, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(cw)
      : "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]));
#else
  mad_row(r, a, b, cw);
#endif
}
template <bool kAlu>
VPIN_HD void mad_row_sel(uint32_t *r, const uint32_t *a, uint32_t b, uint32_t &cw) {
  if (kAlu) mad_row_alu(r, a, b, cw);
  else mad_row(r, a, b, cw);
}
// mul_8x8 with a choice per row of where its accumulation runs: bit 2 i of kAluRows -> the even-column products of row i
// go through mad_row_alu, bit 2 i + 1 -> the odd-column products. kAluRows = 0 is mul_8x8 below.
template <uint32_t kAluRows>
VPIN_HD void mul_8x8_p(uint32_t *t, const uint32_t *a, const uint32_t *b);

// t[0..16) = a * b (8 x 8 limbs). 64 IMAD.WIDE + 14 carry words + one 15-limb add chain.
VPIN_HD void mul_8x8(uint32_t *t, const uint32_t *a, const uint32_t *b) {
  uint32_t ev[16], od[16];  // od[k] has weight 2^(32 (k + 1))
#pragma unroll
  for (int k = 8; k < 16; k++) ev[k] = od[k] = 0;
  mul_row(ev, a, b[0]);
  mul_row(od, a + 1, b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    if (i & 1) {
      mad_row(od + i - 1, a, b[i], od[i + 7]);
      if (i + 9 < 16) mad_row(ev + i + 1, a + 1, b[i], ev[i + 9]);
      else mad_row_nc(ev + i + 1, a + 1, b[i]);
    } else {
      mad_row(ev + i, a, b[i], ev[i + 8]);
      mad_row(od + i, a + 1, b[i], od[i + 8]);
    }
  }
  // t = ev + (od << 32)
  t[0] = ev[0];
#if defined(__CUDA_ARCH__)
  uint32_t cy;
  asm("add.cc.u32 %0, %8, %15; addc.cc.u32 %1, %9, %16; addc.cc.u32 %2, %10, %17; addc.cc.u32 %3, %11, %18;"
      "addc.cc.u32 %4, %12, %19; addc.cc.u32 %5, %13, %20; addc.cc.u32 %6, %14, %21; addc.u32 %7, 0, 0;"
      : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(cy)
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]));
  asm("add.cc.u32 %8, %8, 0xffffffff;"
      "addc.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.u32 %7, %16, %24;"
      : "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "+r"(cy)
      : "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
        "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
#else
  uint64_t c = 0;
  for (int k = 1; k < 16; k++) {
    c += (uint64_t)ev[k] + od[k - 1];
    t[k] = (uint32_t)c;
    c >>= 32;
  }
#endif
}

template <uint32_t kAluRows>
VPIN_HD void mul_8x8_p(uint32_t *t, const uint32_t *a, const uint32_t *b) {
  uint32_t ev[16], od[16];  // od[k] has weight 2^(32 (k + 1))
#pragma unroll
  for (int k = 8; k < 16; k++) ev[k] = od[k] = 0;
  mul_row(ev, a, b[0]);
  mul_row(od, a + 1, b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    const bool alu_e = ((kAluRows >> (2 * i)) & 1u) != 0, alu_o = ((kAluRows >> (2 * i + 1)) & 1u) != 0;
    if (i & 1) {
      if (alu_e) mad_row_alu(od + i - 1, a, b[i], od[i + 7]); else mad_row(od + i - 1, a, b[i], od[i + 7]);
      if (i + 9 < 16) { if (alu_o) mad_row_alu(ev + i + 1, a + 1, b[i], ev[i + 9]); else mad_row(ev + i + 1, a + 1, b[i], ev[i + 9]); }
      else if (alu_o) { uint32_t dummy = 0; mad_row_alu(ev + i + 1, a + 1, b[i], dummy); }
      else mad_row_nc(ev + i + 1, a + 1, b[i]);
    } else {
      if (alu_e) mad_row_alu(ev + i, a, b[i], ev[i + 8]); else mad_row(ev + i, a, b[i], ev[i + 8]);
      if (alu_o) mad_row_alu(od + i, a + 1, b[i], od[i + 8]); else mad_row(od + i, a + 1, b[i], od[i + 8]);
    }
  }
  // t = ev + (od << 32)
  t[0] = ev[0];
#if defined(__CUDA_ARCH__)
  uint32_t cy;
  asm("add.cc.u32 %0, %8, %15; addc.cc.u32 %1, %9, %16; addc.cc.u32 %2, %10, %17; addc.cc.u32 %3, %11, %18;"
      "addc.cc.u32 %4, %12, %19; addc.cc.u32 %5, %13, %20; addc.cc.u32 %6, %14, %21; addc.u32 %7, 0, 0;"
      : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(cy)
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]));
  asm("add.cc.u32 %8, %8, 0xffffffff;"
      "addc.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.u32 %7, %16, %24;"
      : "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "+r"(cy)
      : "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
        "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
#else
  uint64_t c = 0;
  for (int k = 1; k < 16; k++) {
    c += (uint64_t)ev[k] + od[k - 1];
    t[k] = (uint32_t)c;
    c >>= 32;
  }
#endif
}

// r[0..2n) += (a[0], a[2], .., a[2n-2]) * b with one carry chain, cw += carry out (n = 1, 2, 3; n = 4 is mad_row)
VPIN_HD void mad_chain3(uint32_t *r, const uint32_t *a, uint32_t b, uint32_t &cw) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %7, %10, %0; madc.hi.cc.u32 %1, %7, %10, %1;"
      "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
      "madc.lo.cc.u32 %4, %9, %10, %4; madc.hi.cc.u32 %5, %9, %10, %5;"
      "addc.u32 %6, %6, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(cw)
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(b));
#else
  uint64_t c = 0;
  for (int k = 0; k < 3; k++) {
    uint64_t p = (uint64_t)a[2 * k] * b;
    uint64_t lo = (uint64_t)r[2 * k] + (uint32_t)p + c;
    r[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)r[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    r[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  cw += (uint32_t)c;
#endif
}
VPIN_HD void mad_chain2(uint32_t *r, const uint32_t *a, uint32_t b, uint32_t &cw) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1;"
      "madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3;"
      "addc.u32 %4, %4, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(cw)
      : "r"(a[0]), "r"(a[2]), "r"(b));
#else
  uint64_t c = 0;
  for (int k = 0; k < 2; k++) {
    uint64_t p = (uint64_t)a[2 * k] * b;
    uint64_t lo = (uint64_t)r[2 * k] + (uint32_t)p + c;
    r[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)r[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    r[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  cw += (uint32_t)c;
#endif
}
VPIN_HD void mad_chain1(uint32_t *r, const uint32_t *a, uint32_t b, uint32_t &cw) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(cw)
      : "r"(a[0]), "r"(b));
#else
  uint64_t p = (uint64_t)a[0] * b;
  uint64_t lo = (uint64_t)r[0] + (uint32_t)p;
  r[0] = (uint32_t)lo;
  uint64_t hi = (uint64_t)r[1] + (uint32_t)(p >> 32) + (lo >> 32);
  r[1] = (uint32_t)hi;
  cw += (uint32_t)(hi >> 32);
#endif
}

// t[0..16) = a * a. The 28 products a_i a_j (i < j) are accumulated once - row i is two carry chains, the factors a_j with
// j - i odd into the odd accumulator and those with j - i even into the even one, each chain's carry word a limb no later row
// has filled yet - then doubled with funnel shifts and the eight squares a_i^2 (disjoint 64-bit slots) are added:
// 36 IMAD.WIDE instead of the 64 of mul_8x8(t, a, a). The dependent chains that end a commitment (60 doublings of the window
// Horner pass, 254 squarings of the ristretto encoding's inverse square root) are latency-bound on exactly this count.
VPIN_HD void sqr_8x8(uint32_t *t, const uint32_t *a) {
  uint32_t ev[16], od[16];  // od[k] has weight 2^(32 (k + 1))
#pragma unroll
  for (int k = 0; k < 16; k++) ev[k] = od[k] = 0;
  mad_row(od + 0, a + 1, a[0], od[8]);       // a0 * (a1, a3, a5, a7)
  mad_chain3(ev + 2, a + 2, a[0], ev[8]);    // a0 * (a2, a4, a6)
  mad_chain3(od + 2, a + 2, a[1], od[8]);    // a1 * (a2, a4, a6)
  mad_chain3(ev + 4, a + 3, a[1], ev[10]);   // a1 * (a3, a5, a7)
  mad_chain3(od + 4, a + 3, a[2], od[10]);   // a2 * (a3, a5, a7)
  mad_chain2(ev + 6, a + 4, a[2], ev[10]);   // a2 * (a4, a6)
  mad_chain2(od + 6, a + 4, a[3], od[10]);   // a3 * (a4, a6)
  mad_chain2(ev + 8, a + 5, a[3], ev[12]);   // a3 * (a5, a7)
  mad_chain2(od + 8, a + 5, a[4], od[12]);   // a4 * (a5, a7)
  mad_chain1(ev + 10, a + 6, a[4], ev[12]);  // a4 * a6
  mad_chain1(od + 10, a + 6, a[5], od[12]);  // a5 * a6
  mad_chain1(ev + 12, a + 7, a[5], ev[14]);  // a5 * a7
  mad_chain1(od + 12, a + 7, a[6], od[14]);  // a6 * a7
  // s = ev + (od << 32) < 2^511
  uint32_t s[16];
  s[0] = ev[0];
#if defined(__CUDA_ARCH__)
  uint32_t cy;
  asm("add.cc.u32 %0, %8, %15; addc.cc.u32 %1, %9, %16; addc.cc.u32 %2, %10, %17; addc.cc.u32 %3, %11, %18;"
      "addc.cc.u32 %4, %12, %19; addc.cc.u32 %5, %13, %20; addc.cc.u32 %6, %14, %21; addc.u32 %7, 0, 0;"
      : "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(cy)
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]));
  asm("add.cc.u32 %8, %8, 0xffffffff;"
      "addc.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.u32 %7, %16, %24;"
      : "=r"(s[8]), "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]), "+r"(cy)
      : "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
        "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
#else
  {
    uint64_t c = 0;
    for (int k = 1; k < 16; k++) {
      c += (uint64_t)ev[k] + od[k - 1];
      s[k] = (uint32_t)c;
      c >>= 32;
    }
  }
#endif
  // t = 2 s + sum_i a_i^2 2^(64 i)
  uint32_t sh[16], dg[16];
  sh[0] = s[0] << 1;
#pragma unroll
  for (int k = 1; k < 16; k++) sh[k] = (s[k] << 1) | (s[k - 1] >> 31);
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int i = 0; i < 8; i++) asm("mul.lo.u32 %0, %2, %2; mul.hi.u32 %1, %2, %2;" : "=r"(dg[2 * i]), "=r"(dg[2 * i + 1]) : "r"(a[i]));
  uint32_t cz;
  asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24; addc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(cz)
      : "r"(sh[0]), "r"(sh[1]), "r"(sh[2]), "r"(sh[3]), "r"(sh[4]), "r"(sh[5]), "r"(sh[6]), "r"(sh[7]),
        "r"(dg[0]), "r"(dg[1]), "r"(dg[2]), "r"(dg[3]), "r"(dg[4]), "r"(dg[5]), "r"(dg[6]), "r"(dg[7]));
  asm("add.cc.u32 %8, %8, 0xffffffff;"
      "addc.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
      "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.u32 %7, %16, %24;"
      : "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "+r"(cz)
      : "r"(sh[8]), "r"(sh[9]), "r"(sh[10]), "r"(sh[11]), "r"(sh[12]), "r"(sh[13]), "r"(sh[14]), "r"(sh[15]),
        "r"(dg[8]), "r"(dg[9]), "r"(dg[10]), "r"(dg[11]), "r"(dg[12]), "r"(dg[13]), "r"(dg[14]), "r"(dg[15]));
#else
  for (int i = 0; i < 8; i++) {
    uint64_t p = (uint64_t)a[i] * a[i];
    dg[2 * i] = (uint32_t)p;
    dg[2 * i + 1] = (uint32_t)(p >> 32);
  }
  uint64_t c = 0;
  for (int k = 0; k < 16; k++) {
    c += (uint64_t)sh[k] + dg[k];
    t[k] = (uint32_t)c;
    c >>= 32;
  }
#endif
}

// ---- Montgomery reduction rows for l = 2^252 + 27742317777372353535851937790883648493 (limbs P0..P3, 0, 0, 0, 2^28) ----
#define VPIN_L_P0 0x5cf5d3edu
#define VPIN_L_P1 0x5812631au
#define VPIN_L_P2 0xa2f79cd6u
#define VPIN_L_P3 0x14def9deu
#define VPIN_L_P7 0x10000000u
#define VPIN_L_INV32 0x12547e1bu  // -(l^-1) mod 2^32

// w += mergeval (if do_merge); m = w * INV; then, continuing the carry of that addition, the limbs of m * l that sit one
// limb above w: o[0..1] += m P1, o[2..3] += m P3, o[4], o[5] ripple, o[6..7] += m 2^28, cw += carry out.
template <bool kMerge>
VPIN_HD void redc_other(uint32_t &w, uint32_t mergeval, uint32_t &m, uint32_t *o, uint32_t &cw) {
#if defined(__CUDA_ARCH__)
  if (kMerge) {
    // (m * 2^28, the top limb of l, is two shifts and two additions: IMAD.WIDE is the scarce half-rate instruction)
    asm("{ .reg .u32 t0, t1;"
        "add.cc.u32 %0, %0, %11; mul.lo.u32 %1, %0, %12;"
        "shl.b32 t0, %1, 28; shr.u32 t1, %1, 4;"
        "madc.lo.cc.u32 %2, %1, %13, %2; madc.hi.cc.u32 %3, %1, %13, %3;"
        "madc.lo.cc.u32 %4, %1, %14, %4; madc.hi.cc.u32 %5, %1, %14, %5;"
        "addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0;"
        "addc.cc.u32 %8, %8, t0; addc.cc.u32 %9, %9, t1;"
        "addc.u32 %10, %10, 0; }"
        : "+r"(w), "=&r"(m), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(cw)
        : "r"(mergeval), "n"(VPIN_L_INV32), "n"(VPIN_L_P1), "n"(VPIN_L_P3));
  } else {
    asm("{ .reg .u32 t0, t1;"
        "mul.lo.u32 %1, %0, %11;"
        "shl.b32 t0, %1, 28; shr.u32 t1, %1, 4;"
        "mad.lo.cc.u32 %2, %1, %12, %2; madc.hi.cc.u32 %3, %1, %12, %3;"
        "madc.lo.cc.u32 %4, %1, %13, %4; madc.hi.cc.u32 %5, %1, %13, %5;"
        "addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0;"
        "addc.cc.u32 %8, %8, t0; addc.cc.u32 %9, %9, t1;"
        "addc.u32 %10, %10, 0; }"
        : "+r"(w), "=&r"(m), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(cw)
        : "n"(VPIN_L_INV32), "n"(VPIN_L_P1), "n"(VPIN_L_P3));
  }
#else
  uint64_t c = 0;
  if (kMerge) {
    c = (uint64_t)w + mergeval;
    w = (uint32_t)c;
    c >>= 32;
  }
  m = w * VPIN_L_INV32;
  const uint32_t mul[4] = {VPIN_L_P1, VPIN_L_P3, 0u, VPIN_L_P7};
  for (int k = 0; k < 4; k++) {
    uint64_t p = (uint64_t)m * mul[k];
    uint64_t lo = (uint64_t)o[2 * k] + (uint32_t)p + c;
    o[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)o[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    o[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  cw += (uint32_t)c;
#endif
}
// s[0..1] += m P0 (s[0] becomes 0), s[2..3] += m P2, s[4..7] ripple, cw += carry out
VPIN_HD void redc_own(uint32_t *s, uint32_t m, uint32_t &cw) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %9, %10, %0; madc.hi.cc.u32 %1, %9, %10, %1;"
      "madc.lo.cc.u32 %2, %9, %11, %2; madc.hi.cc.u32 %3, %9, %11, %3;"
      "addc.cc.u32 %4, %4, 0; addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0;"
      "addc.u32 %8, %8, 0;"
      : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]), "+r"(s[6]), "+r"(s[7]), "+r"(cw)
      : "r"(m), "n"(VPIN_L_P0), "n"(VPIN_L_P2));
#else
  uint64_t c = 0;
  const uint32_t mul[4] = {VPIN_L_P0, VPIN_L_P2, 0u, 0u};
  for (int k = 0; k < 4; k++) {
    uint64_t p = (uint64_t)m * mul[k];
    uint64_t lo = (uint64_t)s[2 * k] + (uint32_t)p + c;
    s[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)s[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
    s[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  cw += (uint32_t)c;
#endif
}

// r[0..8) = a * b / 2^256 mod l, in [0, 2l): product rows interleaved with reduction rows (word-serial Montgomery),
// all carries confined to 8-limb windows plus one small carry word. 96 IMAD.WIDE (64 product + 32 reduction) + 8 IMAD in total.
VPIN_HD void mont_mul_l(uint32_t *r, const uint32_t *a, const uint32_t *b) {
  uint32_t ev[18], od[18];  // od[k] has weight 2^(32 (k + 1))
#pragma unroll
  for (int k = 8; k < 18; k++) ev[k] = od[k] = 0;
  uint32_t m;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i == 0) {
      mul_row(ev, a, b[0]);
      mul_row(od, a + 1, b[0]);
    } else if (i & 1) {
      mad_row(od + i - 1, a, b[i], od[i + 7]);
      mad_row(ev + i + 1, a + 1, b[i], ev[i + 9]);
    } else {
      mad_row(ev + i, a, b[i], ev[i + 8]);
      mad_row(od + i, a + 1, b[i], od[i + 8]);
    }
    if (i == 0) {
      redc_other<false>(ev[0], 0u, m, od, od[8]);
      redc_own(ev, m, ev[8]);
    } else if (i & 1) {
      redc_other<true>(od[i - 1], ev[i], m, ev + i + 1, ev[i + 9]);
      redc_own(od + i - 1, m, od[i + 7]);
    } else {
      redc_other<true>(ev[i], od[i - 1], m, od + i, od[i + 8]);
      redc_own(ev + i, m, ev[i + 8]);
    }
  }
  // r = ev[8..16) + od[7..15)
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
        "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
#else
  uint64_t c = 0;
  for (int k = 0; k < 8; k++) {
    c += (uint64_t)ev[8 + k] + od[7 + k];
    r[k] = (uint32_t)c;
    c >>= 32;
  }
#endif
}

// r[0..8) = a / 2^256 mod l, in [0, 2l): mont_mul_l(r, a, 1) without its 64 product multiplications - the eight reduction rows
// only (32 IMAD.WIDE + 8 IMAD). The MSM recode kernels leave Montgomery form once per scalar and were throttled by the
// multiply pipe doing it with a full multiplication by one.
VPIN_HD void mont_redc_l(uint32_t *r, const uint32_t *a) {
  uint32_t ev[18], od[18];  // od[k] has weight 2^(32 (k + 1))
#pragma unroll
  for (int k = 0; k < 18; k++) ev[k] = od[k] = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { ev[2 * k] = a[2 * k]; od[2 * k] = a[2 * k + 1]; }
  uint32_t m;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i == 0) {
      redc_other<false>(ev[0], 0u, m, od, od[8]);
      redc_own(ev, m, ev[8]);
    } else if (i & 1) {
      redc_other<true>(od[i - 1], ev[i], m, ev + i + 1, ev[i + 9]);
      redc_own(od + i - 1, m, od[i + 7]);
    } else {
      redc_other<true>(ev[i], od[i - 1], m, od + i, od[i + 8]);
      redc_own(ev + i, m, ev[i + 8]);
    }
  }
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
        "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
#else
  uint64_t c = 0;
  for (int k = 0; k < 8; k++) {
    c += (uint64_t)ev[8 + k] + od[7 + k];
    r[k] = (uint32_t)c;
    c >>= 32;
  }
#endif
}

}  // namespace limb
}  // namespace vpin
