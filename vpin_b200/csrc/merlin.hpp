// Host-side Fiat-Shamir for the B200 prover: Keccak-f[1600], SHAKE256, STROBE-128, Merlin transcripts and
// Spartan's helpers on top. north_star keeps the transcript on the host. Replaces what the reference reaches
// through merlin 3.0.0 / sha3 0.8.2 (Spartan/src/transcript.rs:19-64, Spartan/src/random.rs:14-30,
// Spartan/src/commitments.rs:20-28).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "fl.cuh"

namespace vpin {

class Keccak {
 public:
  // The permutation is compiled twice: plain x86-64, and with BMI1/BMI2 (ANDN for chi, RORX for rho: 0.33 instead of 0.46 - 0.67 us
  // per permutation on the development host) - chosen once at run time. A CNN-A proof absorbs ~0.5 MB through ~4500 permutations
  // on the host's critical path.
  static void permute(uint64_t a[25]) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool bmi = __builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2");
    if (bmi) { permute_bmi(a); return; }
#endif
    permute_body(a);
  }

 private:
#if defined(__x86_64__) && defined(__GNUC__)
  __attribute__((target("bmi,bmi2"))) static void permute_bmi(uint64_t a[25]) { permute_body(a); }
#endif
  __attribute__((always_inline)) static inline void permute_body(uint64_t a[25]) {
    static const uint64_t kRound[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    // lanes in locals, theta / rho+pi / chi written out per plane (the transcript absorbs every commitment row and every
    // opened scalar vector, ~10^4 permutations per proof, so this is on the host critical path)
    uint64_t a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3], a04 = a[4], a05 = a[5], a06 = a[6], a07 = a[7], a08 = a[8], a09 = a[9],
             a10 = a[10], a11 = a[11], a12 = a[12], a13 = a[13], a14 = a[14], a15 = a[15], a16 = a[16], a17 = a[17], a18 = a[18],
             a19 = a[19], a20 = a[20], a21 = a[21], a22 = a[22], a23 = a[23], a24 = a[24];
    for (int rnd = 0; rnd < 24; rnd++) {
      uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
               c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
      uint64_t d0 = c4 ^ rotl(c1, 1), d1 = c0 ^ rotl(c2, 1), d2 = c1 ^ rotl(c3, 1), d3 = c2 ^ rotl(c4, 1), d4 = c3 ^ rotl(c0, 1);
      // b[y + 5 * ((2x + 3y) % 5)] = rotl(a[x + 5y] ^ d[x], rho[x + 5y])
      uint64_t b00 = a00 ^ d0, b10 = rotl(a01 ^ d1, 1), b20 = rotl(a02 ^ d2, 62), b05 = rotl(a03 ^ d3, 28), b15 = rotl(a04 ^ d4, 27);
      uint64_t b16 = rotl(a05 ^ d0, 36), b01 = rotl(a06 ^ d1, 44), b11 = rotl(a07 ^ d2, 6), b21 = rotl(a08 ^ d3, 55), b06 = rotl(a09 ^ d4, 20);
      uint64_t b07 = rotl(a10 ^ d0, 3), b17 = rotl(a11 ^ d1, 10), b02 = rotl(a12 ^ d2, 43), b12 = rotl(a13 ^ d3, 25), b22 = rotl(a14 ^ d4, 39);
      uint64_t b23 = rotl(a15 ^ d0, 41), b08 = rotl(a16 ^ d1, 45), b18 = rotl(a17 ^ d2, 15), b03 = rotl(a18 ^ d3, 21), b13 = rotl(a19 ^ d4, 8);
      uint64_t b14 = rotl(a20 ^ d0, 18), b24 = rotl(a21 ^ d1, 2), b09 = rotl(a22 ^ d2, 61), b19 = rotl(a23 ^ d3, 56), b04 = rotl(a24 ^ d4, 14);
      a00 = b00 ^ (~b01 & b02) ^ kRound[rnd]; a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
      a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
      a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
      a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
      a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
    }
    a[0] = a00; a[1] = a01; a[2] = a02; a[3] = a03; a[4] = a04; a[5] = a05; a[6] = a06; a[7] = a07; a[8] = a08; a[9] = a09;
    a[10] = a10; a[11] = a11; a[12] = a12; a[13] = a13; a[14] = a14; a[15] = a15; a[16] = a16; a[17] = a17; a[18] = a18; a[19] = a19;
    a[20] = a20; a[21] = a21; a[22] = a22; a[23] = a23; a[24] = a24;
  }
  __attribute__((always_inline)) static inline uint64_t rotl(uint64_t v, int n) { return n == 0 ? v : (v << n) | (v >> (64 - n)); }
};

// SHAKE256 extendable-output function (rate 136 bytes)
class Shake256Xof {
 public:
  Shake256Xof() : fill_(0), out_(false) { memset(lanes_, 0, sizeof(lanes_)); }
  void update(const void *data, size_t n) {
    const uint8_t *p = (const uint8_t *)data;
    uint8_t *s = (uint8_t *)lanes_;
    while (n--) {
      s[fill_++] ^= *p++;
      if (fill_ == kRate) { Keccak::permute(lanes_); fill_ = 0; }
    }
  }
  void read(uint8_t *dst, size_t n) {
    uint8_t *s = (uint8_t *)lanes_;
    if (!out_) {
      s[fill_] ^= 0x1f;
      s[kRate - 1] ^= 0x80;
      Keccak::permute(lanes_);
      fill_ = 0;
      out_ = true;
    }
    while (n--) {
      if (fill_ == kRate) { Keccak::permute(lanes_); fill_ = 0; }
      *dst++ = s[fill_++];
    }
  }

 private:
  static const size_t kRate = 136;
  uint64_t lanes_[25];
  size_t fill_;
  bool out_;
};

// STROBE-128/1600 restricted to the operations Merlin uses (meta-AD, AD, PRF)
class Strobe {
 public:
  explicit Strobe(const char *proto) : cursor_(0), op_start_(0) {
    memset(lanes_, 0, sizeof(lanes_));
    uint8_t *s = bytes();
    s[0] = 1; s[1] = kRate + 2; s[2] = 1; s[3] = 0; s[4] = 1; s[5] = 96;
    memcpy(s + 6, "STROBEv1.0.2", 12);
    Keccak::permute(lanes_);
    start(kMeta | kApp);
    mix(proto, strlen(proto));
  }
  void meta(const void *d, size_t n, bool cont) { if (!cont) start(kMeta | kApp); mix(d, n); }
  void data(const void *d, size_t n) { start(kApp); mix(d, n); }
  void prf(uint8_t *out, size_t n) {
    start(kInbound | kApp | kCipher);
    uint8_t *s = bytes();
    for (size_t i = 0; i < n; i++) {
      out[i] = s[cursor_];
      s[cursor_] = 0;
      if (++cursor_ == kRate) flush();
    }
  }

 private:
  enum { kInbound = 1, kApp = 2, kCipher = 4, kMeta = 16, kKey = 32 };
  static const uint8_t kRate = 166;
  uint8_t *bytes() { return (uint8_t *)lanes_; }
  void flush() {
    uint8_t *s = bytes();
    s[cursor_] ^= op_start_;
    s[cursor_ + 1] ^= 0x04;
    s[kRate + 1] ^= 0x80;
    Keccak::permute(lanes_);
    cursor_ = 0;
    op_start_ = 0;
  }
  void mix(const void *d, size_t n) {  // (run by run up to the end of the rate: the inner loop has no branch and vectorises)
    const uint8_t *p = (const uint8_t *)d;
    uint8_t *s = bytes();
    while (n) {
      size_t run = (size_t)kRate - cursor_;
      if (run > n) run = n;
      uint8_t *dst = s + cursor_;
      for (size_t i = 0; i < run; i++) dst[i] ^= p[i];
      p += run;
      n -= run;
      cursor_ = (uint8_t)(cursor_ + run);
      if (cursor_ == kRate) flush();
    }
  }
  void start(uint8_t flags) {
    uint8_t hdr[2] = {op_start_, flags};
    op_start_ = (uint8_t)(cursor_ + 1);
    mix(hdr, 2);
    if ((flags & (kCipher | kKey)) && cursor_ != 0) flush();
  }
  uint64_t lanes_[25];
  uint8_t cursor_, op_start_;
};

// merlin::Transcript with Spartan's ProofTranscript/AppendToTranscript conventions
class MerlinTranscript {
 public:
  MerlinTranscript(const void *label, size_t n) : st_("Merlin v1.0") { message("dom-sep", label, n); }
  explicit MerlinTranscript(const char *label) : MerlinTranscript(label, strlen(label)) {}
  void message(const char *label, const void *msg, size_t n) {
    uint32_t len = (uint32_t)n;
    st_.meta(label, strlen(label), false);
    st_.meta(&len, 4, true);
    st_.data(msg, n);
  }
  void message(const char *label, const char *msg) { message(label, msg, strlen(msg)); }
  void challenge_bytes(const char *label, uint8_t *out, size_t n) {
    uint32_t len = (uint32_t)n;
    st_.meta(label, strlen(label), false);
    st_.meta(&len, 4, true);
    st_.prf(out, n);
  }
  // Spartan/src/transcript.rs:19-43
  void protocol_name(const char *name) { message("protocol-name", name); }
  void scalar(const char *label, const fl_t &x) {
    uint8_t b[32];
    fl_to_bytes(x, b);
    message(label, b, 32);
  }
  void point(const char *label, const uint8_t comp[32]) { message(label, comp, 32); }
  fl_t challenge_scalar(const char *label) {
    uint8_t w[64];
    challenge_bytes(label, w, 64);
    return fl_from_bytes_wide(w);
  }
  std::vector<fl_t> challenge_vector(const char *label, size_t n) {
    std::vector<fl_t> v(n);
    for (size_t i = 0; i < n; i++) v[i] = challenge_scalar(label);
    return v;
  }
  // Spartan/src/transcript.rs:56-64
  void scalars(const char *label, const fl_t *v, size_t n) {
    message(label, "begin_append_vector");
    for (size_t i = 0; i < n; i++) scalar(label, v[i]);
    message(label, "end_append_vector");
  }
  void scalars(const char *label, const std::vector<fl_t> &v) { scalars(label, v.data(), v.size()); }

 private:
  Strobe st_;
};

// Spartan/src/random.rs:7-30; the OsRng scalar is an explicit input (determinism hook of the C ABI)
class ProverTape {
 public:
  ProverTape(const void *name, size_t n, const fl_t &init_randomness) : t_(name, n) { t_.scalar("init_randomness", init_randomness); }
  fl_t scalar(const char *label) { return t_.challenge_scalar(label); }
  std::vector<fl_t> vector(const char *label, size_t n) { return t_.challenge_vector(label, n); }

 private:
  MerlinTranscript t_;
};

}  // namespace vpin
