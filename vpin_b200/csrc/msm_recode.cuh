// Signed-digit recoding of one scalar for the fixed-base MSM tables (kernels_msm.cuh), shared by the bulk recode kernel
// and the fused bullet-reduction kernel.
#pragma once
#include "kernels_msm.cuh"

namespace vpin {

__device__ __forceinline__ uint32_t msm_half_l_limb(int i) {  // (l - 1) / 2
  switch (i) {
    case 0: return 0x2e7ae9f6u; case 1: return 0x2c09318du; case 2: return 0x517bce6bu; case 3: return 0x0a6f7cefu;
    case 7: return 0x08000000u; default: return 0u;
  }
}
template <int W>
__device__ __forceinline__ uint32_t msm_recode_digits(const uint32_t (&v)[9], bool gt, uint16_t *dst, size_t plane, uint32_t *used_out) {
  constexpr int kWindows = 252 / W + 1;  // == msm_geom(W).windows
  constexpr uint32_t kMask = (1u << W) - 1u, kHalf = 1u << (W - 1);
  uint32_t carry = 0, nz = 0, used = 0;
#pragma unroll
  for (int w = 0; w < kWindows; w++) {
    const int bit = w * W, limb = bit >> 5, sh = bit & 31;
    uint32_t raw = (sh == 0 ? v[limb] : __funnelshift_r(v[limb], v[limb + 1], sh)) & kMask;
    raw += carry;
    uint32_t neg = raw > kHalf ? 1u : 0u;
    uint32_t mag = neg ? (1u << W) - raw : raw;
    carry = neg;
    uint32_t sign = (neg ^ (gt ? 1u : 0u)) & (mag != 0 ? 1u : 0u);
    nz += mag != 0 ? 1u : 0u;
    used |= mag != 0 ? (1u << w) : 0u;
    dst[(size_t)w * plane] = (uint16_t)(mag | (sign << 15));
  }
  *used_out = used;
  return nz;
}
// x: Montgomery form. Writes the g.windows digits of the representative of smallest absolute value (|s| <= (l-1)/2) to
// dst[w * plane] as magnitude | sign << 15 and returns the number of non-zero digits; *used_windows (optional) receives the set of
// windows that got a non-zero digit.
__device__ __forceinline__ uint32_t msm_recode_value(const fl_t &x, const MsmGeom &g, uint16_t *dst, size_t plane, uint32_t *used_windows = nullptr) {
  if (used_windows) *used_windows = 0;
  if (fl_is_zero(x)) {
    for (int w = 0; w < g.windows; w++) dst[(size_t)w * plane] = 0;
    return 0;
  }
  fl_t s = fl_from_mont(x);
  // s > (l-1)/2 ?  then use l - s and flip every sign
  bool gt = false;
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    uint32_t h = msm_half_l_limb(i);
    if (s.v[i] != h) { gt = s.v[i] > h; break; }
  }
  uint32_t v[9];
  if (gt) {
    int64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { int64_t t = (int64_t)fl_modulus_limb(i) - (int64_t)s.v[i] + br; v[i] = (uint32_t)t; br = t >> 32; }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = s.v[i];
  }
  v[8] = 0;
  uint32_t nz = 0, used = 0;
  // one unrolled copy of the digit loop per window width: window w starts at the compile-time bit w * W, so the funnel over
  // (v[limb], v[limb + 1]) is one shift on fixed registers (the run-time loop spent ~40 instructions per window on select
  // chains, which made the recode kernels instruction-bound at 0.2 of the HBM rate)
  switch (g.W) {
    case 12: nz = msm_recode_digits<12>(v, gt, dst, plane, &used); break;
    case 13: nz = msm_recode_digits<13>(v, gt, dst, plane, &used); break;
    case 14: nz = msm_recode_digits<14>(v, gt, dst, plane, &used); break;
    default: nz = msm_recode_digits<15>(v, gt, dst, plane, &used); break;
  }
  if (used_windows) *used_windows = used;
  return nz;
}

}  // namespace vpin
