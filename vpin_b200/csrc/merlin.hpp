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
  static void permute(uint64_t a[25]) {
    static const uint64_t kRound[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    // rho offsets indexed [x + 5*y]
    static const int kRho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int rnd = 0; rnd < 24; rnd++) {
      uint64_t c[5], d[5], b[25];
      for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
      for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
      for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
      for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(a[x + 5 * y], kRho[x + 5 * y]);
      for (int y = 0; y < 5; y++)
        for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
      a[0] ^= kRound[rnd];
    }
  }

 private:
  static uint64_t rotl(uint64_t v, int n) { return n == 0 ? v : (v << n) | (v >> (64 - n)); }
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
  void mix(const void *d, size_t n) {
    const uint8_t *p = (const uint8_t *)d;
    uint8_t *s = bytes();
    for (size_t i = 0; i < n; i++) {
      s[cursor_] ^= p[i];
      if (++cursor_ == kRate) flush();
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
