// CPU test (tests/test_host_transcript.py): the product's host transcript (vpin_b200/csrc/merlin.hpp - STROBE run-wise XOR,
// unrolled Keccak) against the oracle's independent restatement (oracle/transcript.hpp, pinned by the merlin crate's test
// vector) on random operation sequences whose messages cross the 166-byte rate boundary at every offset.
#include "../../vpin_b200/csrc/merlin.hpp"
#include "../../oracle/transcript.hpp"
#include <cstdio>
#include <random>
#include <vector>
int main() {
  std::mt19937_64 rng(7);
  int bad = 0;
  for (int trial = 0; trial < 300; trial++) {
    vpin::MerlinTranscript a("parity");
    orc::Transcript b("parity");
    for (int op = 0; op < 40; op++) {
      size_t n = (rng() % 5 == 0) ? rng() % 1200 : rng() % 70;
      std::vector<uint8_t> m(n);
      for (auto &x : m) x = (uint8_t)rng();
      if (rng() % 4 == 0) {
        size_t k = 1 + rng() % 400;
        std::vector<uint8_t> o1(k), o2(k);
        a.challenge_bytes("chal", o1.data(), k);
        b.challenge_bytes("chal", o2.data(), k);
        if (o1 != o2) bad++;
      } else {
        a.message("lbl", m.data(), n);
        b.append_message("lbl", m.data(), n);
      }
    }
    uint8_t o1[64], o2[64];
    a.challenge_bytes("fin", o1, 64);
    b.challenge_bytes("fin", o2, 64);
    if (memcmp(o1, o2, 64)) bad++;
  }
  printf("mismatches=%d\n", bad);
  return bad != 0;
}
