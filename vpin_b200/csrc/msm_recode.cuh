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
  uint32_t carry = 0, nz = 0, used = 0;
  // limb indexing by a run-time window: v lives in local memory for this loop only when the compiler cannot resolve it;
  // the funnel over (v[limb], v[limb + 1]) is written with a select chain to stay in registers
  const uint32_t wmask = (1u << g.W) - 1u;
  for (int w = 0; w < g.windows; w++) {
    int bit = w * g.W, limb = bit >> 5, sh = bit & 31;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { lo = limb == i ? v[i] : lo; hi = limb == i ? v[i + 1] : hi; }
    uint64_t two = (uint64_t)lo | ((uint64_t)hi << 32);
    uint32_t raw = ((uint32_t)(two >> sh) & wmask) + carry;
    uint32_t neg = raw > (uint32_t)g.table ? 1u : 0u;
    uint32_t mag = neg ? (1u << g.W) - raw : raw;
    carry = neg;
    uint32_t sign = (neg ^ (gt ? 1u : 0u)) & (mag != 0 ? 1u : 0u);
    nz += mag != 0 ? 1u : 0u;
    used |= mag != 0 ? (1u << w) : 0u;
    dst[(size_t)w * plane] = (uint16_t)(mag | (sign << 15));
  }
  if (used_windows) *used_windows = used;
  return nz;
}

}  // namespace vpin
