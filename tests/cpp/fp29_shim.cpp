// Host build of vpin_b200/csrc/fp29.cuh (the radix-2^29 F_p of the MSM hot loop) for tests/test_fp29.py.
#include "../../vpin_b200/csrc/fp29.cuh"
using namespace vpin;
extern "C" {
void fp29_mul(const int32_t *a, const int32_t *b, int unsigned_variant, int32_t *out) {
  f9 x, y;
  for (int k = 0; k < 9; k++) { x.v[k] = a[k]; y.v[k] = b[k]; }
  f9 r = unsigned_variant ? f9_mul<true>(x, y) : f9_mul<false>(x, y);
  for (int k = 0; k < 9; k++) out[k] = r.v[k];
}
void fp29_unpack(const uint32_t *w, int32_t *out) {
  f9 r = f9_unpack(w);
  for (int k = 0; k < 9; k++) out[k] = r.v[k];
}
void fp29_to_fp(const int32_t *a, uint32_t *out) {
  f9 x;
  for (int k = 0; k < 9; k++) x.v[k] = a[k];
  fp_t r = f9_to_fp(x);
  for (int k = 0; k < 8; k++) out[k] = r.v[k];
}
}
