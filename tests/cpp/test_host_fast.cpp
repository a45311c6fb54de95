// CPU check of the prover's host-side 64-bit arithmetic (vpin_b200/csrc/host_fast.hpp, fl_mul_host64 in fl.cuh) against
// the oracle's independent field / group code. Test infrastructure: built and run by tests/test_host_fast.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../oracle/ed.hpp"
#include "../../vpin_b200/csrc/host_fast.hpp"

static std::mt19937_64 rng(0x7650494E);
static int fails = 0;
#define CHECK(c, msg) do { if (!(c)) { printf("FAIL: %s (line %d)\n", msg, __LINE__); fails++; } } while (0)

static orc::Fl rand_fl() {
  uint8_t w[64];
  for (int i = 0; i < 64; i++) w[i] = (uint8_t)rng();
  return orc::fl_from_bytes_wide(w);
}
static vpin::fl_t to_v(const orc::Fl &a) { vpin::fl_t r; memcpy(r.v, a.v, 32); return r; }
static bool same(const vpin::fl_t &a, const orc::Fl &b) { return memcmp(a.v, b.v, 32) == 0; }
static vpin::ge_t to_dev(const orc::Pt &p) {
  uint8_t b[32];
  vpin::ge_t g;
  orc::fp_tobytes(p.X, b); g.X = vpin::fp_from_bytes(b);
  orc::fp_tobytes(p.Y, b); g.Y = vpin::fp_from_bytes(b);
  orc::fp_tobytes(p.Z, b); g.Z = vpin::fp_from_bytes(b);
  orc::fp_tobytes(p.T, b); g.T = vpin::fp_from_bytes(b);
  return g;
}

int main() {
  // ---- F_l: host Montgomery multiplication ----
  orc::Fl edge[5] = {orc::fl_zero(), orc::fl_one(), orc::fl_neg(orc::fl_one()), orc::FL_R2, orc::Fl{{1, 0, 0, 0}}};
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) CHECK(same(vpin::fl_mul(to_v(edge[i]), to_v(edge[j])), orc::fl_mul(edge[i], edge[j])), "fl_mul edge");
  for (int it = 0; it < 20000; it++) {
    orc::Fl a = rand_fl(), b = rand_fl();
    CHECK(same(vpin::fl_mul(to_v(a), to_v(b)), orc::fl_mul(a, b)), "fl_mul");
    CHECK(same(vpin::fl_add(to_v(a), to_v(b)), orc::fl_add(a, b)), "fl_add");
    CHECK(same(vpin::fl_sub(to_v(a), to_v(b)), orc::fl_sub(a, b)), "fl_sub");
  }
  {
    orc::Fl a = rand_fl();
    CHECK(same(vpin::fl_invert(to_v(a)), orc::fl_invert(a)), "fl_invert");
    uint8_t w[64];
    for (int i = 0; i < 64; i++) w[i] = 0xff;
    CHECK(same(vpin::fl_from_bytes_wide(w), orc::fl_from_bytes_wide(w)), "from_bytes_wide 0xff");
  }
  // ---- F_p: lazily reduced device limbs -> 5 x 51 ----
  {
    vpin::fp_t all1;
    for (int i = 0; i < 8; i++) all1.v[i] = 0xffffffffu;  // 2^256 - 1 == 37 (mod p)
    uint8_t b[32], want[32] = {37};
    vpin::hf::fe_to_bytes(vpin::hf::fe_from_fp(all1), b);
    CHECK(memcmp(b, want, 32) == 0, "fe_from_fp(2^256-1)");
    vpin::fp_t pm1 = all1;  // p - 1 + p = 2p - 1 -> canonical p - 1 ... use p + 5 instead: limbs of 2^255 - 19 + 5
    pm1.v[0] = 0xfffffff2u; pm1.v[7] = 0x7fffffffu;
    uint8_t want5[32] = {5};
    vpin::hf::fe_to_bytes(vpin::hf::fe_from_fp(pm1), b);
    CHECK(memcmp(b, want5, 32) == 0, "fe_from_fp(p+5)");
  }
  // ---- group: fixed-base multiplication, accumulation, encoding ----
  orc::Pt B;
  CHECK(orc::pt_decompress(orc::BASEPOINT_COMPRESSED, &B), "basepoint");
  orc::Pt bases[3];
  vpin::hf::FixedBase fb[3];
  for (int k = 0; k < 3; k++) {
    bases[k] = orc::pt_mul(rand_fl(), B);
    for (int d = 0; d < k; d++) bases[k] = orc::pt_double(bases[k]);  // non-trivial Z
    fb[k].build(vpin::hf::ge_from_dev(to_dev(bases[k])));
  }
  for (int it = 0; it < 300; it++) {
    orc::Fl s[3] = {rand_fl(), rand_fl(), rand_fl()};
    if (it == 0) s[0] = orc::fl_zero();
    if (it == 1) s[0] = orc::fl_neg(orc::fl_one());
    if (it == 2) s[1] = orc::fl_one();
    if (it == 3) { uint64_t v[4] = {0x8080808080808080ull, 0x8080808080808080ull, 0x8080808080808080ull, 0x0080808080808080ull}; s[2] = orc::fl_from_raw(v); }
    if (it == 4) { uint64_t v[4] = {0x8181818181818181ull, 0x7f7f7f7f7f7f7f7full, 0x80ff80ff80ff80ffull, 0x0fffffffffffffffull}; s[2] = orc::fl_from_raw(v); }
    vpin::hf::ge acc = vpin::hf::ge_identity();
    for (int k = 0; k < 3; k++) fb[k].mul_acc(to_v(s[k]), &acc);
    orc::Pt want = orc::msm(s, bases, 3);
    uint8_t a[32], w[32];
    vpin::hf::ge_compress(acc, a);
    orc::pt_compress(want, w);
    CHECK(memcmp(a, w, 32) == 0, "fixed-base msm + compress");
    // also the one-shot mul and the identity encoding
    vpin::hf::ge_compress(fb[0].mul(to_v(s[0])), a);
    orc::pt_compress(orc::pt_mul(s[0], bases[0]), w);
    CHECK(memcmp(a, w, 32) == 0, "fixed-base mul");
  }
  // host full addition / doubling against the oracle
  {
    vpin::hf::ge p = vpin::hf::ge_from_dev(to_dev(bases[1])), q = vpin::hf::ge_from_dev(to_dev(bases[2]));
    uint8_t a[32], w[32];
    vpin::hf::ge_compress(vpin::hf::ge_add(p, q), a);
    orc::pt_compress(orc::pt_add(bases[1], bases[2]), w);
    CHECK(memcmp(a, w, 32) == 0, "ge_add");
    vpin::hf::ge_compress(vpin::hf::ge_dbl(p), a);
    orc::pt_compress(orc::pt_double(bases[1]), w);
    CHECK(memcmp(a, w, 32) == 0, "ge_dbl");
  }
  printf(fails ? "host_fast: %d failures\n" : "host_fast: ok\n", fails);
  return fails ? 1 : 0;
}
