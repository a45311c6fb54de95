// CPU test (tests/test_host_transcript.py): the portable (host) paths of the limb routines added in round 2 against the
// routines they specialise - sqr_8x8 against mul_8x8(a, a), mont_redc_l against mont_mul_l(a, 1). The device paths are the same
// index structure with the carry chains written in PTX; they are exercised bit for bit by every GPU parity test (a commitment
// cannot be encoded without the squaring, a scalar not recoded without the reduction).
#include "../../vpin_b200/csrc/limbs.cuh"
#include <cstdio>
#include <cstring>
#include <random>
using namespace vpin::limb;
int main() {
  std::mt19937_64 rng(1);
  int bad = 0;
  for (int it = 0; it < 200000; it++) {
    uint32_t a[8], t1[16], t2[16], one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, r1[8], r2[8];
    for (int i = 0; i < 8; i++) {
      uint64_t r = rng();
      a[i] = (uint32_t)r;
      int m = (r >> 32) % 8;
      if (it % 5 == 1 && m < 3) a[i] = 0xffffffffu;
      if (it % 5 == 2 && m < 3) a[i] = 0;
      if (it % 7 == 3) a[i] = 0xffffffffu;
    }
    mul_8x8(t1, a, a);
    sqr_8x8(t2, a);
    if (memcmp(t1, t2, 64)) bad++;
    if (it % 3 == 0) a[7] &= 0x1fffffffu;  // (values below 2 l, as the kernels hold them)
    mont_mul_l(r1, a, one);
    mont_redc_l(r2, a);
    if (memcmp(r1, r2, 32)) bad++;
  }
  printf("mismatches=%d\n", bad);
  return bad != 0;
}
