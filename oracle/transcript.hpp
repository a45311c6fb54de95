// TEST INFRASTRUCTURE ONLY (CPU oracle). Not linked into the product library.
//
// Keccak-f[1600], SHAKE256, STROBE-128 and the Merlin transcript. These live in crates that are NOT vendored
// under /root/reference: merlin 3.0.0 (-> keccak 0.1.4) and sha3 0.8.2 (vPIN_proof_generation/Cargo.lock).
// Restated from FIPS 202, the STROBE v1.0.2 spec and the Merlin spec; pinned by the Merlin crate's published
// "test protocol" vector and hashlib.shake_256 in tests/test_oracle_primitives.py.
// Spartan's layer on top follows Spartan/src/transcript.rs:19-43 and Spartan/src/random.rs:14-30.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "ed.hpp"

namespace orc {

static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static inline void keccak_f1600(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
      0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
      0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
      0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
  static const int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
  uint64_t bc[5], t;
  for (int round = 0; round < 24; round++) {
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; i++) {
      t = bc[(i + 4) % 5] ^ rol64(bc[(i + 1) % 5], 1);
      for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
    }
    t = st[1];
    for (int i = 0; i < 24; i++) {
      int j = PILN[i];
      bc[0] = st[j];
      st[j] = rol64(t, ROTC[i]);
      t = bc[0];
    }
    for (int j = 0; j < 25; j += 5) {
      for (int i = 0; i < 5; i++) bc[i] = st[j + i];
      for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
    }
    st[0] ^= RC[round];
  }
}

// SHAKE256 XOF (FIPS 202): rate 136, domain suffix 0x1f
struct Shake256 {
  uint64_t st[25];
  size_t pos;
  bool squeezing;
  Shake256() : pos(0), squeezing(false) { memset(st, 0, sizeof(st)); }
  void absorb(const uint8_t *data, size_t n) {
    uint8_t *b = (uint8_t *)st;
    for (size_t i = 0; i < n; i++) {
      b[pos++] ^= data[i];
      if (pos == 136) { keccak_f1600(st); pos = 0; }
    }
  }
  void squeeze(uint8_t *out, size_t n) {
    uint8_t *b = (uint8_t *)st;
    if (!squeezing) {
      b[pos] ^= 0x1f;
      b[135] ^= 0x80;
      keccak_f1600(st);
      pos = 0;
      squeezing = true;
    }
    for (size_t i = 0; i < n; i++) {
      if (pos == 136) { keccak_f1600(st); pos = 0; }
      out[i] = b[pos++];
    }
  }
};

struct Strobe128 {
  static const int R = 166;
  enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
  uint64_t st64[25];
  uint8_t pos, pos_begin, cur_flags;
  uint8_t *st() { return (uint8_t *)st64; }
  explicit Strobe128(const char *protocol_label) {
    memset(st64, 0, sizeof(st64));
    const uint8_t hdr[6] = {1, R + 2, 1, 0, 1, 96};
    memcpy(st(), hdr, 6);
    memcpy(st() + 6, "STROBEv1.0.2", 12);
    keccak_f1600(st64);
    pos = 0; pos_begin = 0; cur_flags = 0;
    meta_ad((const uint8_t *)protocol_label, strlen(protocol_label), false);
  }
  void run_f() {
    st()[pos] ^= pos_begin;
    st()[pos + 1] ^= 0x04;
    st()[R + 1] ^= 0x80;
    keccak_f1600(st64);
    pos = 0; pos_begin = 0;
  }
  void absorb(const uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      st()[pos] ^= d[i];
      pos++;
      if (pos == R) run_f();
    }
  }
  void squeeze(uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      d[i] = st()[pos];
      st()[pos] = 0;
      pos++;
      if (pos == R) run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) return;  // continuing the current operation
    uint8_t old_begin = pos_begin;
    pos_begin = pos + 1;
    cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    bool force_f = (flags & (FLAG_C | FLAG_K)) != 0;
    if (force_f && pos != 0) run_f();
  }
  void meta_ad(const uint8_t *d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
  void ad(const uint8_t *d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
  void prf(uint8_t *d, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(d, n); }
};

// merlin::Transcript + Spartan's ProofTranscript / AppendToTranscript (Spartan/src/transcript.rs)
struct Transcript {
  Strobe128 s;
  explicit Transcript(const char *label, size_t label_len) : s("Merlin v1.0") {
    append_message("dom-sep", (const uint8_t *)label, label_len);
  }
  explicit Transcript(const char *label) : Transcript(label, strlen(label)) {}
  void append_message(const char *label, const uint8_t *msg, size_t n) {
    uint32_t len = (uint32_t)n;
    s.meta_ad((const uint8_t *)label, strlen(label), false);
    s.meta_ad((const uint8_t *)&len, 4, true);
    s.ad(msg, n, false);
  }
  void append_message(const char *label, const char *msg) { append_message(label, (const uint8_t *)msg, strlen(msg)); }
  void append_u64(const char *label, uint64_t x) { append_message(label, (const uint8_t *)&x, 8); }
  void challenge_bytes(const char *label, uint8_t *out, size_t n) {
    uint32_t len = (uint32_t)n;
    s.meta_ad((const uint8_t *)label, strlen(label), false);
    s.meta_ad((const uint8_t *)&len, 4, true);
    s.prf(out, n, false);
  }
  // transcript.rs:19-43
  void append_protocol_name(const char *name) { append_message("protocol-name", name); }
  void append_scalar(const char *label, const Fl &x) {
    uint8_t b[32];
    fl_to_bytes(x, b);
    append_message(label, b, 32);
  }
  void append_point(const char *label, const uint8_t comp[32]) { append_message(label, comp, 32); }
  Fl challenge_scalar(const char *label) {
    uint8_t buf[64];
    challenge_bytes(label, buf, 64);
    return fl_from_bytes_wide(buf);
  }
  std::vector<Fl> challenge_vector(const char *label, size_t n) {
    std::vector<Fl> v(n);
    for (size_t i = 0; i < n; i++) v[i] = challenge_scalar(label);
    return v;
  }
  // transcript.rs:56-64 (AppendToTranscript for [Scalar])
  void append_scalars(const char *label, const std::vector<Fl> &v) {
    append_message(label, "begin_append_vector");
    for (const Fl &x : v) append_scalar(label, x);
    append_message(label, "end_append_vector");
  }
};

// Spartan/src/random.rs:14-30 with the OsRng-drawn scalar supplied by the caller (determinism hook)
struct RandomTape {
  Transcript tape;
  RandomTape(const char *name, size_t name_len, const Fl &init_randomness) : tape(name, name_len) {
    tape.append_scalar("init_randomness", init_randomness);
  }
  Fl random_scalar(const char *label) { return tape.challenge_scalar(label); }
  std::vector<Fl> random_vector(const char *label, size_t n) { return tape.challenge_vector(label, n); }
};

}  // namespace orc
