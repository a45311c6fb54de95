// TEST INFRASTRUCTURE ONLY (CPU oracle). Not linked into the product library.
//
// CPU restatement of the Spartan SNARK prover AND verifier exactly as vPIN drives them.
// "SP/" = /root/reference/src/proof_generation/Spartan/src/, "VP/" = .../vPIN_proof_generation/src/.
// Every function cites the reference lines it follows. The reference seeds its RandomTape from OsRng
// (SP/random.rs:16-18); here the 32-byte init_randomness scalar is an explicit argument.
// Parity status: F_l, UniPoly, MLE and eq-table order are pinned by the reference's own KATs; the group,
// transcript and bincode layers are pinned by RFC 9496 / libsodium / the Merlin vector (dalek, merlin and
// bincode are not vendored in the reference and no reference test fixes a proof byte) — see DESIGN.md.
#pragma once
#include <array>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <thread>

#include "transcript.hpp"

namespace orc {

typedef std::vector<Fl> FlVec;
typedef std::array<uint8_t, 32> Comp;

static inline Fl operator+(const Fl &a, const Fl &b) { return fl_add(a, b); }
static inline Fl operator-(const Fl &a, const Fl &b) { return fl_sub(a, b); }
static inline Fl operator*(const Fl &a, const Fl &b) { return fl_mul(a, b); }
static inline Fl &operator+=(Fl &a, const Fl &b) { a = fl_add(a, b); return a; }
static inline Fl &operator-=(Fl &a, const Fl &b) { a = fl_sub(a, b); return a; }
static inline Fl &operator*=(Fl &a, const Fl &b) { a = fl_mul(a, b); return a; }

static int g_threads = 1;  // threads for Hyrax rows only (SP/dense_mlpoly.rs:160-175 rayon); everything else 1 thread

// SP/math.rs:29-35
static inline size_t log_2(size_t x) {
  assert(x != 0);
  size_t fl = 63 - __builtin_clzll((unsigned long long)x);
  return (x & (x - 1)) == 0 ? fl : fl + 1;
}
static inline size_t pow2(size_t e) { return (size_t)1 << e; }
static inline size_t next_pow2(size_t x) { size_t p = 1; while (p < x) p <<= 1; return p; }

static inline Comp compress(const Pt &p) { Comp c; pt_compress(p, c.data()); return c; }
static inline Pt decompress_or_die(const Comp &c) {
  Pt p;
  if (!pt_decompress(c.data(), &p)) throw std::runtime_error("decompress failed");
  return p;
}

// ----------------------------------------------------------------------------------------------
// SP/commitments.rs
struct MultiCommitGens {
  size_t n;
  std::vector<Pt> G;
  Pt h;
};
// SP/commitments.rs:20-38
static inline MultiCommitGens mcg_new(size_t n, const char *label) {
  Shake256 sh;
  sh.absorb((const uint8_t *)label, strlen(label));
  sh.absorb(BASEPOINT_COMPRESSED, 32);
  std::vector<Pt> gens(n + 1);
  uint8_t buf[64];
  for (size_t i = 0; i < n + 1; i++) { sh.squeeze(buf, 64); gens[i] = pt_from_uniform_bytes(buf); }
  MultiCommitGens g;
  g.n = n;
  g.h = gens[n];
  gens.resize(n);
  g.G = std::move(gens);
  return g;
}
// SP/commitments.rs:57-72
static inline void mcg_split_at(const MultiCommitGens &g, size_t mid, MultiCommitGens *a, MultiCommitGens *b) {
  a->n = mid; a->G.assign(g.G.begin(), g.G.begin() + mid); a->h = g.h;
  b->n = g.G.size() - mid; b->G.assign(g.G.begin() + mid, g.G.end()); b->h = g.h;
}
// SP/commitments.rs:79-84
static inline Pt commit_scalar(const Fl &x, const Fl &blind, const MultiCommitGens &g) {
  assert(g.n == 1);
  Fl s[2] = {x, blind};
  Pt p[2] = {g.G[0], g.h};
  return msm(s, p, 2);
}
// SP/commitments.rs:86-98
static inline Pt commit_vec(const Fl *v, size_t n, const Fl &blind, const MultiCommitGens &g) {
  assert(g.n == n);
  return pt_add(msm(v, g.G.data(), n), pt_mul(blind, g.h));
}

// SP/nizk/mod.rs:411-425
struct DotProductProofGens { size_t n; MultiCommitGens gens_n, gens_1; };
static inline DotProductProofGens dppg_new(size_t n, const char *label) {
  DotProductProofGens g;
  g.n = n;
  mcg_split_at(mcg_new(n + 1, label), n, &g.gens_n, &g.gens_1);
  return g;
}
// SP/dense_mlpoly.rs:28-41, :96-98
static inline void factored_lens(size_t ell, size_t *l, size_t *r) { *l = ell / 2; *r = ell - ell / 2; }
struct PolyCommitmentGens { DotProductProofGens gens; };
static inline PolyCommitmentGens pcg_new(size_t num_vars, const char *label) {
  size_t l, r;
  factored_lens(num_vars, &l, &r);
  return PolyCommitmentGens{dppg_new(pow2(r), label)};
}
// SP/r1csproof.rs:49-90
struct R1CSSumcheckGens { MultiCommitGens gens_1, gens_3, gens_4; };
struct R1CSGens { R1CSSumcheckGens gens_sc; PolyCommitmentGens gens_pc; };
static inline R1CSGens r1csgens_new(const char *label, size_t num_vars) {
  R1CSGens g;
  g.gens_pc = pcg_new(log_2(num_vars), label);
  g.gens_sc.gens_1 = g.gens_pc.gens.gens_1;
  g.gens_sc.gens_3 = mcg_new(3, label);
  g.gens_sc.gens_4 = mcg_new(4, label);
  return g;
}
// SP/sparse_mlpoly.rs:294-328
struct SparseMatPolyCommitmentGens { PolyCommitmentGens gens_ops, gens_mem, gens_derefs; };
static inline SparseMatPolyCommitmentGens smpcg_new(const char *label, size_t num_vars_x, size_t num_vars_y,
                                                    size_t num_nz_entries, size_t batch_size) {
  size_t num_vars_ops = log_2(next_pow2(num_nz_entries)) + log_2(next_pow2(batch_size * 5));
  size_t num_vars_mem = (num_vars_x > num_vars_y ? num_vars_x : num_vars_y) + 1;
  size_t num_vars_derefs = log_2(next_pow2(num_nz_entries)) + log_2(next_pow2(batch_size * 2));
  return SparseMatPolyCommitmentGens{pcg_new(num_vars_ops, label), pcg_new(num_vars_mem, label),
                                     pcg_new(num_vars_derefs, label)};
}
// SP/lib.rs:295-327, SP/r1csinstance.rs:33-48
struct SNARKGens { R1CSGens gens_r1cs_sat; SparseMatPolyCommitmentGens gens_r1cs_eval; };
static inline SNARKGens snarkgens_new(size_t num_cons, size_t num_vars, size_t num_inputs, size_t num_nz_entries) {
  size_t num_vars_padded = num_vars > num_inputs + 1 ? num_vars : num_inputs + 1;
  num_vars_padded = next_pow2(num_vars_padded);
  SNARKGens g;
  g.gens_r1cs_sat = r1csgens_new("gens_r1cs_sat", num_vars_padded);
  assert(num_inputs < num_vars_padded);
  g.gens_r1cs_eval = smpcg_new("gens_r1cs_eval", log_2(num_cons), log_2(2 * num_vars_padded), num_nz_entries, 3);
  return g;
}

// ----------------------------------------------------------------------------------------------
// bincode 1.3.3 default config (fixed-width LE ints, Vec = u64 len + items, structs/tuples/arrays concatenated)
struct Ar {
  bool writing;
  std::vector<uint8_t> *out;
  const uint8_t *in;
  size_t pos, len;
  bool ok;
  void raw(void *p, size_t n) {
    if (writing) { out->insert(out->end(), (uint8_t *)p, (uint8_t *)p + n); return; }
    if (pos + n > len) { ok = false; memset(p, 0, n); return; }
    memcpy(p, in + pos, n);
    pos += n;
  }
};
static inline void io(Ar &a, uint64_t &x) { a.raw(&x, 8); }
static inline void io(Ar &a, Fl &x) { a.raw(x.v, 32); }  // raw Montgomery limbs (SP/scalar/ristretto255.rs:199-200)
static inline void io(Ar &a, Comp &c) { a.raw(c.data(), 32); }
template <class T>
static inline void io(Ar &a, std::vector<T> &v) {
  uint64_t n = v.size();
  io(a, n);
  if (!a.writing) {
    if (n > (a.len - a.pos)) { a.ok = false; return; }
    v.resize(n);
  }
  for (auto &x : v) io(a, x);
}

// ----------------------------------------------------------------------------------------------
// SP/dense_mlpoly.rs
// :78-94 (r[0] is the most significant index bit)
static inline FlVec eq_evals(const FlVec &r) {
  size_t ell = r.size();
  FlVec evals(pow2(ell), fl_one());
  size_t size = 1;
  for (size_t j = 0; j < ell; j++) {
    size *= 2;
    for (size_t i = size; i-- > 0;) {
      if ((i & 1) == 0) continue;  // i runs over size-1, size-3, ...
      Fl scalar = evals[i / 2];
      evals[i] = scalar * r[j];
      evals[i - 1] = scalar - evals[i];
    }
  }
  return evals;
}
// :71-76
static inline Fl eq_evaluate(const FlVec &r, const FlVec &rx) {
  assert(r.size() == rx.size());
  Fl acc = fl_one(), one = fl_one();
  for (size_t i = 0; i < rx.size(); i++) acc = acc * (r[i] * rx[i] + (one - r[i]) * (one - rx[i]));
  return acc;
}
// :100-108
static inline void eq_factored_evals(const FlVec &r, FlVec *L, FlVec *R) {
  size_t l, rr;
  factored_lens(r.size(), &l, &rr);
  *L = eq_evals(FlVec(r.begin(), r.begin() + l));
  *R = eq_evals(FlVec(r.begin() + l, r.end()));
}
// :121-127
static inline Fl identity_poly_evaluate(const FlVec &r) {
  size_t len = r.size();
  Fl acc = fl_zero();
  for (size_t i = 0; i < len; i++) acc += fl_from_u64((uint64_t)pow2(len - i - 1)) * r[i];
  return acc;
}

struct DensePoly {
  size_t num_vars, len;
  FlVec Z;
  DensePoly() : num_vars(0), len(0) {}
  explicit DensePoly(FlVec z) : num_vars(log_2(z.size())), len(z.size()), Z(std::move(z)) {}
  const Fl &operator[](size_t i) const { return Z[i]; }
  // :229-236
  void bound_poly_var_top(const Fl &r) {
    size_t n = len / 2;
    for (size_t i = 0; i < n; i++) Z[i] = Z[i] + r * (Z[i + n] - Z[i]);
    num_vars -= 1;
    len = n;
  }
  // :238-245
  void bound_poly_var_bot(const Fl &r) {
    size_t n = len / 2;
    for (size_t i = 0; i < n; i++) Z[i] = Z[2 * i] + r * (Z[2 * i + 1] - Z[2 * i]);
    num_vars -= 1;
    len = n;
  }
};
static inline Fl dotproduct(const Fl *a, const Fl *b, size_t n) {
  Fl acc = fl_zero();
  for (size_t i = 0; i < n; i++) acc += a[i] * b[i];
  return acc;
}
// :249-255
static inline Fl dense_evaluate(const DensePoly &p, const FlVec &r) {
  assert(r.size() == p.num_vars);
  FlVec chis = eq_evals(r);
  assert(chis.size() == p.Z.size());
  return dotproduct(p.Z.data(), chis.data(), chis.size());
}
// :220-227
static inline FlVec dense_bound(const DensePoly &p, const FlVec &L) {
  size_t l, r;
  factored_lens(p.num_vars, &l, &r);
  size_t L_size = pow2(l), R_size = pow2(r);
  FlVec out(R_size, fl_zero());
  for (size_t j = 0; j < L_size; j++) {
    const Fl *row = &p.Z[j * R_size];
    if (fl_is_zero(L[j])) continue;
    for (size_t i = 0; i < R_size; i++) out[i] += L[j] * row[i];
  }
  return out;
}
// :272-285
static inline DensePoly dense_merge(const std::vector<const DensePoly *> &polys) {
  FlVec Z;
  for (const DensePoly *p : polys) Z.insert(Z.end(), p->Z.begin(), p->Z.end());
  Z.resize(next_pow2(Z.size()), fl_zero());
  return DensePoly(std::move(Z));
}
static inline DensePoly dense_from_usize(const std::vector<size_t> &z) {
  FlVec Z(z.size());
  for (size_t i = 0; i < z.size(); i++) Z[i] = fl_from_u64((uint64_t)z[i]);
  return DensePoly(std::move(Z));
}

struct PolyCommitment { std::vector<Comp> C; };
static inline void io(Ar &a, PolyCommitment &c) { io(a, c.C); }
// :305-313
static inline void append_poly_commitment(Transcript &t, const char *label, const PolyCommitment &c) {
  t.append_message(label, "poly_commitment_begin");
  for (const Comp &p : c.C) t.append_point("poly_commitment_share", p.data());
  t.append_message(label, "poly_commitment_end");
}
// :160-175 (rows in parallel, as the reference's rayon path)
static inline PolyCommitment commit_inner(const FlVec &Z, const FlVec &blinds, const MultiCommitGens &gens) {
  size_t L_size = blinds.size(), R_size = Z.size() / L_size;
  assert(L_size * R_size == Z.size());
  PolyCommitment pc;
  pc.C.resize(L_size);
  auto work = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) pc.C[i] = compress(commit_vec(&Z[R_size * i], R_size, blinds[i], gens));
  };
  int nt = g_threads;
  if (nt <= 1 || L_size < 2) { work(0, L_size); return pc; }
  std::vector<std::thread> th;
  size_t chunk = (L_size + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    size_t lo = t * chunk, hi = lo + chunk < L_size ? lo + chunk : L_size;
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto &x : th) x.join();
  return pc;
}
// :193-218
static inline PolyCommitment dense_commit(const DensePoly &p, const PolyCommitmentGens &gens, RandomTape *tape,
                                          FlVec *blinds_out) {
  size_t l, r;
  factored_lens(p.num_vars, &l, &r);
  size_t L_size = pow2(l);
  FlVec blinds = tape ? tape->random_vector("poly_blinds", L_size) : FlVec(L_size, fl_zero());
  PolyCommitment c = commit_inner(p.Z, blinds, gens.gens.gens_n);
  if (blinds_out) *blinds_out = blinds;
  return c;
}
// VP/commit_test.rs:27-57
static inline PolyCommitment my_dense_mlpoly_commit(const DensePoly &p, const PolyCommitmentGens &gens,
                                                    const FlVec &blind_1, const FlVec &blind_2, FlVec *blinds_out) {
  assert(blind_1.size() == blind_2.size());
  FlVec sum(blind_1.size());
  for (size_t i = 0; i < sum.size(); i++) sum[i] = blind_1[i] + blind_2[i];
  PolyCommitment c = commit_inner(p.Z, sum, gens.gens.gens_n);
  *blinds_out = sum;
  return c;
}

// ----------------------------------------------------------------------------------------------
// SP/unipoly.rs
struct UniPoly {
  FlVec coeffs;
  // :23-54
  static UniPoly from_evals(const FlVec &e) {
    assert(e.size() == 3 || e.size() == 4);
    UniPoly p;
    Fl one = fl_one();
    Fl two_inv = fl_invert(one + one);
    if (e.size() == 3) {
      Fl c = e[0];
      Fl a = two_inv * (e[2] - e[1] - e[1] + c);
      Fl b = e[1] - c - a;
      p.coeffs = {c, b, a};
    } else {
      Fl six_inv = fl_invert(one + one + one + one + one + one);
      Fl d = e[0];
      Fl a = six_inv * (e[3] - e[2] - e[2] - e[2] + e[1] + e[1] + e[1] - e[0]);
      Fl b = two_inv * (e[0] + e[0] - e[1] - e[1] - e[1] - e[1] - e[1] + e[2] + e[2] + e[2] + e[2] - e[3]);
      Fl c = e[1] - d - a - b;
      p.coeffs = {d, c, b, a};
    }
    return p;
  }
  size_t degree() const { return coeffs.size() - 1; }
  Fl eval_at_zero() const { return coeffs[0]; }
  Fl eval_at_one() const { Fl s = fl_zero(); for (const Fl &c : coeffs) s += c; return s; }
  // :70-78
  Fl evaluate(const Fl &r) const {
    Fl eval = coeffs[0], power = r;
    for (size_t i = 1; i < coeffs.size(); i++) { eval += power * coeffs[i]; power *= r; }
    return eval;
  }
  // :80-87
  FlVec compress() const {
    FlVec c;
    c.push_back(coeffs[0]);
    c.insert(c.end(), coeffs.begin() + 2, coeffs.end());
    return c;
  }
  // :96-108
  static UniPoly decompress(const FlVec &cel, const Fl &hint) {
    Fl linear = hint - cel[0] - cel[0];
    for (size_t i = 1; i < cel.size(); i++) linear -= cel[i];
    UniPoly p;
    p.coeffs = {cel[0], linear};
    p.coeffs.insert(p.coeffs.end(), cel.begin() + 1, cel.end());
    return p;
  }
  // :111-119
  void append_to_transcript(Transcript &t, const char *label) const {
    t.append_message(label, "UniPoly_begin");
    for (const Fl &c : coeffs) t.append_scalar("coeff", c);
    t.append_message(label, "UniPoly_end");
  }
};

// ----------------------------------------------------------------------------------------------
// SP/nizk/mod.rs
struct KnowledgeProof { Comp alpha; Fl z1, z2; };
static inline void io(Ar &a, KnowledgeProof &p) { io(a, p.alpha); io(a, p.z1); io(a, p.z2); }
// :28-53
static inline KnowledgeProof knowledge_prove(const MultiCommitGens &g, Transcript &t, RandomTape &tape, const Fl &x,
                                             const Fl &r, Comp *C_out) {
  t.append_protocol_name("knowledge proof");
  Fl t1 = tape.random_scalar("t1"), t2 = tape.random_scalar("t2");
  Comp C = compress(commit_scalar(x, r, g));
  t.append_point("C", C.data());
  Comp alpha = compress(commit_scalar(t1, t2, g));
  t.append_point("alpha", alpha.data());
  Fl c = t.challenge_scalar("c");
  *C_out = C;
  return KnowledgeProof{alpha, x * c + t1, r * c + t2};
}
// :55-75
static inline bool knowledge_verify(const KnowledgeProof &p, const MultiCommitGens &g, Transcript &t, const Comp &C) {
  t.append_protocol_name("knowledge proof");
  t.append_point("C", C.data());
  t.append_point("alpha", p.alpha.data());
  Fl c = t.challenge_scalar("c");
  Pt Cp, Ap;
  if (!pt_decompress(C.data(), &Cp) || !pt_decompress(p.alpha.data(), &Ap)) return false;
  Comp lhs = compress(commit_scalar(p.z1, p.z2, g));
  Comp rhs = compress(pt_add(pt_mul(c, Cp), Ap));
  return lhs == rhs;
}

struct EqualityProof { Comp alpha; Fl z; };
static inline void io(Ar &a, EqualityProof &p) { io(a, p.alpha); io(a, p.z); }
// :90-118
static inline EqualityProof equality_prove(const MultiCommitGens &g, Transcript &t, RandomTape &tape, const Fl &v1,
                                           const Fl &s1, const Fl &v2, const Fl &s2) {
  t.append_protocol_name("equality proof");
  Fl r = tape.random_scalar("r");
  Comp C1 = compress(commit_scalar(v1, s1, g));
  t.append_point("C1", C1.data());
  Comp C2 = compress(commit_scalar(v2, s2, g));
  t.append_point("C2", C2.data());
  Comp alpha = compress(pt_mul(r, g.h));
  t.append_point("alpha", alpha.data());
  Fl c = t.challenge_scalar("c");
  return EqualityProof{alpha, c * (s1 - s2) + r};
}
// :120-145
static inline bool equality_verify(const EqualityProof &p, const MultiCommitGens &g, Transcript &t, const Comp &C1,
                                   const Comp &C2) {
  t.append_protocol_name("equality proof");
  t.append_point("C1", C1.data());
  t.append_point("C2", C2.data());
  t.append_point("alpha", p.alpha.data());
  Fl c = t.challenge_scalar("c");
  Pt P1, P2, A;
  if (!pt_decompress(C1.data(), &P1) || !pt_decompress(C2.data(), &P2) || !pt_decompress(p.alpha.data(), &A)) return false;
  Comp rhs = compress(pt_add(pt_mul(c, pt_sub(P1, P2)), A));
  Comp lhs = compress(pt_mul(p.z, g.h));
  return lhs == rhs;
}

struct ProductProof { Comp alpha, beta, delta; Fl z[5]; };
static inline void io(Ar &a, ProductProof &p) {
  io(a, p.alpha); io(a, p.beta); io(a, p.delta);
  for (int i = 0; i < 5; i++) io(a, p.z[i]);
}
// :162-232
static inline ProductProof product_prove(const MultiCommitGens &g, Transcript &t, RandomTape &tape, const Fl &x,
                                         const Fl &rX, const Fl &y, const Fl &rY, const Fl &z, const Fl &rZ, Comp *Xo,
                                         Comp *Yo, Comp *Zo) {
  t.append_protocol_name("product proof");
  Fl b1 = tape.random_scalar("b1"), b2 = tape.random_scalar("b2"), b3 = tape.random_scalar("b3"),
     b4 = tape.random_scalar("b4"), b5 = tape.random_scalar("b5");
  Comp X = compress(commit_scalar(x, rX, g));
  t.append_point("X", X.data());
  Comp Y = compress(commit_scalar(y, rY, g));
  t.append_point("Y", Y.data());
  Comp Z = compress(commit_scalar(z, rZ, g));
  t.append_point("Z", Z.data());
  Comp alpha = compress(commit_scalar(b1, b2, g));
  t.append_point("alpha", alpha.data());
  Comp beta = compress(commit_scalar(b3, b4, g));
  t.append_point("beta", beta.data());
  MultiCommitGens gX;
  gX.n = 1; gX.G = {decompress_or_die(X)}; gX.h = g.h;
  Comp delta = compress(commit_scalar(b3, b5, gX));
  t.append_point("delta", delta.data());
  Fl c = t.challenge_scalar("c");
  ProductProof p;
  p.alpha = alpha; p.beta = beta; p.delta = delta;
  p.z[0] = b1 + c * x;
  p.z[1] = b2 + c * rX;
  p.z[2] = b3 + c * y;
  p.z[3] = b4 + c * rY;
  p.z[4] = b5 + c * (rZ - rX * y);
  *Xo = X; *Yo = Y; *Zo = Z;
  return p;
}
// :234-246
static inline bool product_check_equality(const Comp &P, const Comp &X, const Fl &c, const MultiCommitGens &g,
                                          const Fl &z1, const Fl &z2) {
  Pt Pp, Xp;
  if (!pt_decompress(P.data(), &Pp) || !pt_decompress(X.data(), &Xp)) return false;
  Comp lhs = compress(pt_add(Pp, pt_mul(c, Xp)));
  Comp rhs = compress(commit_scalar(z1, z2, g));
  return lhs == rhs;
}
// :248-292
static inline bool product_verify(const ProductProof &p, const MultiCommitGens &g, Transcript &t, const Comp &X,
                                  const Comp &Y, const Comp &Z) {
  t.append_protocol_name("product proof");
  t.append_point("X", X.data());
  t.append_point("Y", Y.data());
  t.append_point("Z", Z.data());
  t.append_point("alpha", p.alpha.data());
  t.append_point("beta", p.beta.data());
  t.append_point("delta", p.delta.data());
  Fl c = t.challenge_scalar("c");
  Pt Xp;
  if (!pt_decompress(X.data(), &Xp)) return false;
  MultiCommitGens gX;
  gX.n = 1; gX.G = {Xp}; gX.h = g.h;
  return product_check_equality(p.alpha, X, c, g, p.z[0], p.z[1]) &&
         product_check_equality(p.beta, Y, c, g, p.z[2], p.z[3]) &&
         product_check_equality(p.delta, Z, c, gX, p.z[2], p.z[4]);
}

struct DotProductProof { Comp delta, beta; FlVec z; Fl z_delta, z_beta; };
static inline void io(Ar &a, DotProductProof &p) { io(a, p.delta); io(a, p.beta); io(a, p.z); io(a, p.z_delta); io(a, p.z_beta); }
// :315-374
static inline DotProductProof dotproduct_prove(const MultiCommitGens &gens_1, const MultiCommitGens &gens_n,
                                               Transcript &t, RandomTape &tape, const FlVec &x_vec, const Fl &blind_x,
                                               const FlVec &a_vec, const Fl &y, const Fl &blind_y) {
  t.append_protocol_name("dot product proof");
  size_t n = x_vec.size();
  assert(a_vec.size() == n && gens_n.n == n && gens_1.n == 1);
  FlVec d_vec = tape.random_vector("d_vec", n);
  Fl r_delta = tape.random_scalar("r_delta"), r_beta = tape.random_scalar("r_beta");
  Comp Cx = compress(commit_vec(x_vec.data(), n, blind_x, gens_n));
  t.append_point("Cx", Cx.data());
  Comp Cy = compress(commit_scalar(y, blind_y, gens_1));
  t.append_point("Cy", Cy.data());
  t.append_scalars("a", a_vec);
  Comp delta = compress(commit_vec(d_vec.data(), n, r_delta, gens_n));
  t.append_point("delta", delta.data());
  Fl dotproduct_a_d = dotproduct(a_vec.data(), d_vec.data(), n);
  Comp beta = compress(commit_scalar(dotproduct_a_d, r_beta, gens_1));
  t.append_point("beta", beta.data());
  Fl c = t.challenge_scalar("c");
  DotProductProof p;
  p.delta = delta; p.beta = beta;
  p.z.resize(n);
  for (size_t i = 0; i < n; i++) p.z[i] = c * x_vec[i] + d_vec[i];
  p.z_delta = c * blind_x + r_delta;
  p.z_beta = c * blind_y + r_beta;
  return p;
}
// :376-408
static inline bool dotproduct_verify(const DotProductProof &p, const MultiCommitGens &gens_1,
                                     const MultiCommitGens &gens_n, Transcript &t, const FlVec &a, const Comp &Cx,
                                     const Comp &Cy) {
  if (gens_n.n != a.size() || gens_1.n != 1 || p.z.size() != a.size()) return false;
  t.append_protocol_name("dot product proof");
  t.append_point("Cx", Cx.data());
  t.append_point("Cy", Cy.data());
  t.append_scalars("a", a);
  t.append_point("delta", p.delta.data());
  t.append_point("beta", p.beta.data());
  Fl c = t.challenge_scalar("c");
  Pt Cxp, Cyp, dp, bp;
  if (!pt_decompress(Cx.data(), &Cxp) || !pt_decompress(Cy.data(), &Cyp) || !pt_decompress(p.delta.data(), &dp) ||
      !pt_decompress(p.beta.data(), &bp))
    return false;
  bool result = pt_eq(pt_add(pt_mul(c, Cxp), dp), commit_vec(p.z.data(), p.z.size(), p.z_delta, gens_n));
  Fl dotproduct_z_a = dotproduct(p.z.data(), a.data(), a.size());
  result &= pt_eq(pt_add(pt_mul(c, Cyp), bp), commit_scalar(dotproduct_z_a, p.z_beta, gens_1));
  return result;
}

// SP/nizk/bullet.rs
struct BulletReductionProof { std::vector<Comp> L_vec, R_vec; };
static inline void io(Ar &a, BulletReductionProof &p) { io(a, p.L_vec); io(a, p.R_vec); }
// bullet.rs:32-132
static inline BulletReductionProof bullet_prove(Transcript &t, const Pt &Q, const std::vector<Pt> &G_vec, const Pt &H,
                                                const FlVec &a_vec, const FlVec &b_vec, const Fl &blind,
                                                const std::vector<std::pair<Fl, Fl>> &blinds_vec, Pt *Gamma_hat,
                                                Fl *a_hat, Fl *b_hat, Pt *g_hat, Fl *blind_fin_out) {
  std::vector<Pt> G = G_vec;
  FlVec a = a_vec, b = b_vec;
  size_t n = G.size();
  assert((n & (n - 1)) == 0);
  size_t lg_n = log_2(n);
  assert(a.size() == n && b.size() == n && blinds_vec.size() == 2 * lg_n);
  BulletReductionProof proof;
  size_t blinds_it = 0;
  Fl blind_fin = blind;
  while (n != 1) {
    n /= 2;
    Fl c_L = dotproduct(&a[0], &b[n], n);  // <a_L, b_R>
    Fl c_R = dotproduct(&a[n], &b[0], n);  // <a_R, b_L>
    const Fl &blind_L = blinds_vec[blinds_it].first, &blind_R = blinds_vec[blinds_it].second;
    blinds_it++;
    FlVec sc(n + 2);
    std::vector<Pt> pts(n + 2);
    for (size_t i = 0; i < n; i++) { sc[i] = a[i]; pts[i] = G[n + i]; }
    sc[n] = c_L; pts[n] = Q; sc[n + 1] = blind_L; pts[n + 1] = H;
    Pt L = msm(sc.data(), pts.data(), n + 2);
    for (size_t i = 0; i < n; i++) { sc[i] = a[n + i]; pts[i] = G[i]; }
    sc[n] = c_R; sc[n + 1] = blind_R;
    Pt R = msm(sc.data(), pts.data(), n + 2);
    Comp Lc = compress(L), Rc = compress(R);
    t.append_point("L", Lc.data());
    t.append_point("R", Rc.data());
    Fl u = t.challenge_scalar("u");
    Fl u_inv = fl_invert(u);
    for (size_t i = 0; i < n; i++) {
      a[i] = a[i] * u + u_inv * a[n + i];
      b[i] = b[i] * u_inv + u * b[n + i];
      Fl s2[2] = {u_inv, u};
      Pt p2[2] = {G[i], G[n + i]};
      G[i] = msm(s2, p2, 2);
    }
    blind_fin = blind_fin + blind_L * u * u + blind_R * u_inv * u_inv;
    proof.L_vec.push_back(Lc);
    proof.R_vec.push_back(Rc);
  }
  Fl s3[3] = {a[0], a[0] * b[0], blind_fin};
  Pt p3[3] = {G[0], Q, H};
  *Gamma_hat = msm(s3, p3, 3);
  *a_hat = a[0]; *b_hat = b[0]; *g_hat = G[0]; *blind_fin_out = blind_fin;
  return proof;
}
// bullet.rs:137-225
static inline bool bullet_verify(const BulletReductionProof &p, size_t n, const FlVec &a, Transcript &t,
                                 const Pt &Gamma, const std::vector<Pt> &G, Pt *G_hat, Pt *Gamma_hat, Fl *a_hat) {
  size_t lg_n = p.L_vec.size();
  if (lg_n >= 32 || n != ((size_t)1 << lg_n) || p.R_vec.size() != lg_n) return false;
  FlVec challenges(lg_n);
  for (size_t i = 0; i < lg_n; i++) {
    t.append_point("L", p.L_vec[i].data());
    t.append_point("R", p.R_vec[i].data());
    challenges[i] = t.challenge_scalar("u");
  }
  FlVec challenges_inv = challenges;
  Fl allinv = fl_batch_invert(challenges_inv);
  for (size_t i = 0; i < lg_n; i++) { challenges[i] = fl_sqr(challenges[i]); challenges_inv[i] = fl_sqr(challenges_inv[i]); }
  FlVec s(n);
  s[0] = allinv;
  for (size_t i = 1; i < n; i++) {
    size_t lg_i = 31 - __builtin_clz((uint32_t)i);
    size_t k = (size_t)1 << lg_i;
    s[i] = s[i - k] * challenges[(lg_n - 1) - lg_i];
  }
  std::vector<Pt> pts(2 * lg_n + 1);
  FlVec sc(2 * lg_n + 1);
  for (size_t i = 0; i < lg_n; i++) {
    if (!pt_decompress(p.L_vec[i].data(), &pts[i]) || !pt_decompress(p.R_vec[i].data(), &pts[lg_n + i])) return false;
    sc[i] = challenges[i];
    sc[lg_n + i] = challenges_inv[i];
  }
  sc[2 * lg_n] = fl_one();
  pts[2 * lg_n] = Gamma;
  *G_hat = msm(s.data(), G.data(), n);
  *a_hat = dotproduct(a.data(), s.data(), n);
  *Gamma_hat = msm(sc.data(), pts.data(), 2 * lg_n + 1);
  return true;
}

struct DotProductProofLog { BulletReductionProof bullet; Comp delta, beta; Fl z1, z2; };
static inline void io(Ar &a, DotProductProofLog &p) { io(a, p.bullet); io(a, p.delta); io(a, p.beta); io(a, p.z1); io(a, p.z2); }
// mod.rs:447-531
static inline DotProductProofLog dotproductlog_prove(const DotProductProofGens &gens, Transcript &t, RandomTape &tape,
                                                     const FlVec &x_vec, const Fl &blind_x, const FlVec &a_vec,
                                                     const Fl &y, const Fl &blind_y, Comp *Cx_out, Comp *Cy_out) {
  t.append_protocol_name("dot product proof (log)");
  size_t n = x_vec.size();
  assert(a_vec.size() == n && gens.n == n);
  Fl d = tape.random_scalar("d");
  Fl r_delta = tape.random_scalar("r_delta");
  Fl r_beta = tape.random_scalar("r_delta");  // sic: the reference reuses the label (mod.rs:466)
  size_t lg = log_2(n);
  FlVec v1 = tape.random_vector("blinds_vec_1", 2 * lg);
  FlVec v2 = tape.random_vector("blinds_vec_2", 2 * lg);
  std::vector<std::pair<Fl, Fl>> blinds_vec(v1.size());
  for (size_t i = 0; i < v1.size(); i++) blinds_vec[i] = {v1[i], v2[i]};
  Comp Cx = compress(commit_vec(x_vec.data(), n, blind_x, gens.gens_n));
  t.append_point("Cx", Cx.data());
  Comp Cy = compress(commit_scalar(y, blind_y, gens.gens_1));
  t.append_point("Cy", Cy.data());
  t.append_scalars("a", a_vec);
  Fl r = t.challenge_scalar("r");
  Pt G1_scaled = pt_mul(r, gens.gens_1.G[0]);  // gens_1.scale(&r), commitments.rs:49-55
  Fl blind_Gamma = blind_x + r * blind_y;
  Pt Gamma_hat, g_hat;
  Fl x_hat, a_hat, rhat_Gamma;
  DotProductProofLog p;
  p.bullet = bullet_prove(t, G1_scaled, gens.gens_n.G, gens.gens_n.h, x_vec, a_vec, blind_Gamma, blinds_vec,
                          &Gamma_hat, &x_hat, &a_hat, &g_hat, &rhat_Gamma);
  Fl y_hat = x_hat * a_hat;
  MultiCommitGens g_hat_gens;
  g_hat_gens.n = 1; g_hat_gens.G = {g_hat}; g_hat_gens.h = gens.gens_1.h;
  p.delta = compress(commit_scalar(d, r_delta, g_hat_gens));
  t.append_point("delta", p.delta.data());
  MultiCommitGens g1s;
  g1s.n = 1; g1s.G = {G1_scaled}; g1s.h = gens.gens_1.h;
  p.beta = compress(commit_scalar(d, r_beta, g1s));
  t.append_point("beta", p.beta.data());
  Fl c = t.challenge_scalar("c");
  p.z1 = d + c * y_hat;
  p.z2 = a_hat * (c * rhat_Gamma + r_beta) + r_delta;
  if (Cx_out) *Cx_out = Cx;
  if (Cy_out) *Cy_out = Cy;
  return p;
}
// mod.rs:533-583
static inline bool dotproductlog_verify(const DotProductProofLog &p, size_t n, const DotProductProofGens &gens,
                                        Transcript &t, const FlVec &a, const Comp &Cx, const Comp &Cy) {
  if (gens.n != n || a.size() != n) return false;
  t.append_protocol_name("dot product proof (log)");
  t.append_point("Cx", Cx.data());
  t.append_point("Cy", Cy.data());
  t.append_scalars("a", a);
  Fl r = t.challenge_scalar("r");
  Pt G1_scaled = pt_mul(r, gens.gens_1.G[0]);
  Pt Cxp, Cyp;
  if (!pt_decompress(Cx.data(), &Cxp) || !pt_decompress(Cy.data(), &Cyp)) return false;
  Pt Gamma = pt_add(Cxp, pt_mul(r, Cyp));
  Pt g_hat, Gamma_hat;
  Fl a_hat;
  if (!bullet_verify(p.bullet, n, a, t, Gamma, gens.gens_n.G, &g_hat, &Gamma_hat, &a_hat)) return false;
  t.append_point("delta", p.delta.data());
  t.append_point("beta", p.beta.data());
  Fl c = t.challenge_scalar("c");
  Pt beta_s, delta_s;
  if (!pt_decompress(p.beta.data(), &beta_s) || !pt_decompress(p.delta.data(), &delta_s)) return false;
  Pt lhs = pt_add(pt_mul(a_hat, pt_add(pt_mul(c, Gamma_hat), beta_s)), delta_s);
  Pt rhs = pt_add(pt_mul(p.z1, pt_add(g_hat, pt_mul(a_hat, G1_scaled))), pt_mul(p.z2, gens.gens_1.h));
  return compress(lhs) == compress(rhs);
}

// SP/dense_mlpoly.rs:315-418
struct PolyEvalProof { DotProductProofLog proof; };
static inline void io(Ar &a, PolyEvalProof &p) { io(a, p.proof); }
// :326-379
static inline PolyEvalProof polyeval_prove(const DensePoly &poly, const FlVec *blinds_opt, const FlVec &r, const Fl &Zr,
                                           const Fl *blind_Zr_opt, const PolyCommitmentGens &gens, Transcript &t,
                                           RandomTape &tape, Comp *C_Zr_prime) {
  t.append_protocol_name("polynomial evaluation proof");
  assert(poly.num_vars == r.size());
  size_t l, rr;
  factored_lens(r.size(), &l, &rr);
  size_t L_size = pow2(l), R_size = pow2(rr);
  FlVec default_blinds(L_size, fl_zero());
  const FlVec &blinds = blinds_opt ? *blinds_opt : default_blinds;
  assert(blinds.size() == L_size);
  Fl blind_Zr = blind_Zr_opt ? *blind_Zr_opt : fl_zero();
  FlVec L, R;
  eq_factored_evals(r, &L, &R);
  assert(L.size() == L_size && R.size() == R_size);
  FlVec LZ = dense_bound(poly, L);
  Fl LZ_blind = dotproduct(blinds.data(), L.data(), L_size);
  PolyEvalProof p;
  p.proof = dotproductlog_prove(gens.gens, t, tape, LZ, LZ_blind, R, Zr, blind_Zr, nullptr, C_Zr_prime);
  return p;
}
// :381-403
static inline bool polyeval_verify(const PolyEvalProof &p, const PolyCommitmentGens &gens, Transcript &t,
                                   const FlVec &r, const Comp &C_Zr, const PolyCommitment &comm) {
  t.append_protocol_name("polynomial evaluation proof");
  FlVec L, R;
  eq_factored_evals(r, &L, &R);
  if (comm.C.size() != L.size()) return false;
  std::vector<Pt> C(comm.C.size());
  for (size_t i = 0; i < C.size(); i++)
    if (!pt_decompress(comm.C[i].data(), &C[i])) return false;
  Comp C_LZ = compress(msm(L.data(), C.data(), L.size()));
  return dotproductlog_verify(p.proof, R.size(), gens.gens, t, R, C_LZ, C_Zr);
}
// :405-417
static inline bool polyeval_verify_plain(const PolyEvalProof &p, const PolyCommitmentGens &gens, Transcript &t,
                                         const FlVec &r, const Fl &Zr, const PolyCommitment &comm) {
  Comp C_Zr = compress(commit_scalar(Zr, fl_zero(), gens.gens.gens_1));
  return polyeval_verify(p, gens, t, r, C_Zr, comm);
}

// ----------------------------------------------------------------------------------------------
// SP/sumcheck.rs
struct SumcheckInstanceProof { std::vector<FlVec> compressed_polys; };  // CompressedUniPoly = Vec<Scalar>
static inline void io(Ar &a, SumcheckInstanceProof &p) { io(a, p.compressed_polys); }
// :27-61
static inline bool sumcheck_verify(const SumcheckInstanceProof &p, const Fl &claim, size_t num_rounds,
                                   size_t degree_bound, Transcript &t, Fl *e_out, FlVec *r_out) {
  Fl e = claim;
  FlVec r;
  if (p.compressed_polys.size() != num_rounds) return false;
  for (size_t i = 0; i < num_rounds; i++) {
    if (p.compressed_polys[i].size() != degree_bound) return false;
    UniPoly poly = UniPoly::decompress(p.compressed_polys[i], e);
    if (poly.degree() != degree_bound) return false;
    if (poly.eval_at_zero() + poly.eval_at_one() != e) return false;
    poly.append_to_transcript(t, "poly");
    Fl r_i = t.challenge_scalar("challenge_nextround");
    r.push_back(r_i);
    e = poly.evaluate(r_i);
  }
  *e_out = e;
  *r_out = r;
  return true;
}

struct ZKSumcheckInstanceProof { std::vector<Comp> comm_polys, comm_evals; std::vector<DotProductProof> proofs; };
static inline void io(Ar &a, ZKSumcheckInstanceProof &p) { io(a, p.comm_polys); io(a, p.comm_evals); io(a, p.proofs); }

// the per-round tail shared by prove_quad (:488-577) and prove_cubic_with_additive_term (:678-767)
static inline void zk_round_tail(const UniPoly &poly, const Fl &r_j, size_t j, const Fl &blind_claim,
                                 const FlVec &blinds_poly, const FlVec &blinds_evals, Fl &claim_per_round,
                                 Comp &comm_claim_per_round, const MultiCommitGens &gens_1,
                                 const MultiCommitGens &gens_n, Transcript &t, RandomTape &tape,
                                 ZKSumcheckInstanceProof &out) {
  Fl eval = poly.evaluate(r_j);
  Comp comm_eval = compress(commit_scalar(eval, blinds_evals[j], gens_1));
  t.append_point("comm_claim_per_round", comm_claim_per_round.data());
  t.append_point("comm_eval", comm_eval.data());
  FlVec w = t.challenge_vector("combine_two_claims_to_one", 2);
  Fl target = w[0] * claim_per_round + w[1] * eval;
  Pt pts[2] = {decompress_or_die(comm_claim_per_round), decompress_or_die(comm_eval)};
  Comp comm_target = compress(msm(w.data(), pts, 2));
  const Fl &blind_sc = j == 0 ? blind_claim : blinds_evals[j - 1];
  Fl blind = w[0] * blind_sc + w[1] * blinds_evals[j];
  if (!(compress(commit_scalar(target, blind, gens_1)) == comm_target)) throw std::runtime_error("sumcheck: comm_target mismatch");
  size_t deg = poly.degree();
  FlVec a_sc(deg + 1, fl_one());
  a_sc[0] += fl_one();
  FlVec a_eval(deg + 1, fl_one());
  for (size_t k = 1; k < a_eval.size(); k++) a_eval[k] = a_eval[k - 1] * r_j;
  FlVec a(deg + 1);
  for (size_t i = 0; i < a.size(); i++) a[i] = w[0] * a_sc[i] + w[1] * a_eval[i];
  DotProductProof proof = dotproduct_prove(gens_1, gens_n, t, tape, poly.coeffs, blinds_poly[j], a, target, blind);
  claim_per_round = eval;
  comm_claim_per_round = comm_eval;
  out.proofs.push_back(proof);
  out.comm_evals.push_back(comm_eval);
}

// :588-776 with comb_func = A*(B*C - D) from SP/r1csproof.rs:104-108
static inline ZKSumcheckInstanceProof zk_prove_cubic_with_additive_term(
    const Fl &claim, const Fl &blind_claim, size_t num_rounds, DensePoly &A, DensePoly &B, DensePoly &C, DensePoly &D,
    const MultiCommitGens &gens_1, const MultiCommitGens &gens_n, Transcript &t, RandomTape &tape, FlVec *r_out,
    FlVec *claims_out, Fl *blind_post) {
  FlVec blinds_poly = tape.random_vector("blinds_poly", num_rounds);
  FlVec blinds_evals = tape.random_vector("blinds_evals", num_rounds);
  Fl claim_per_round = claim;
  Comp comm_claim_per_round = compress(commit_scalar(claim_per_round, blind_claim, gens_1));
  ZKSumcheckInstanceProof out;
  FlVec r;
  auto comb = [](const Fl &a, const Fl &b, const Fl &c, const Fl &d) { return a * (b * c - d); };
  for (size_t j = 0; j < num_rounds; j++) {
    Fl e0 = fl_zero(), e2 = fl_zero(), e3 = fl_zero();
    size_t len = A.len / 2;
    for (size_t i = 0; i < len; i++) {
      e0 += comb(A[i], B[i], C[i], D[i]);
      Fl a2 = A[len + i] + A[len + i] - A[i], b2 = B[len + i] + B[len + i] - B[i];
      Fl c2 = C[len + i] + C[len + i] - C[i], d2 = D[len + i] + D[len + i] - D[i];
      e2 += comb(a2, b2, c2, d2);
      Fl a3 = a2 + A[len + i] - A[i], b3 = b2 + B[len + i] - B[i];
      Fl c3 = c2 + C[len + i] - C[i], d3 = d2 + D[len + i] - D[i];
      e3 += comb(a3, b3, c3, d3);
    }
    UniPoly poly = UniPoly::from_evals({e0, claim_per_round - e0, e2, e3});
    Comp comm_poly = compress(commit_vec(poly.coeffs.data(), poly.coeffs.size(), blinds_poly[j], gens_n));
    t.append_point("comm_poly", comm_poly.data());
    out.comm_polys.push_back(comm_poly);
    Fl r_j = t.challenge_scalar("challenge_nextround");
    A.bound_poly_var_top(r_j);
    B.bound_poly_var_top(r_j);
    C.bound_poly_var_top(r_j);
    D.bound_poly_var_top(r_j);
    zk_round_tail(poly, r_j, j, blind_claim, blinds_poly, blinds_evals, claim_per_round, comm_claim_per_round, gens_1,
                  gens_n, t, tape, out);
    r.push_back(r_j);
  }
  *r_out = r;
  *claims_out = {A[0], B[0], C[0], D[0]};
  *blind_post = blinds_evals[num_rounds - 1];
  return out;
}
// :428-586 with comb_func = A*B from SP/r1csproof.rs:139-140
static inline ZKSumcheckInstanceProof zk_prove_quad(const Fl &claim, const Fl &blind_claim, size_t num_rounds,
                                                    DensePoly &A, DensePoly &B, const MultiCommitGens &gens_1,
                                                    const MultiCommitGens &gens_n, Transcript &t, RandomTape &tape,
                                                    FlVec *r_out, FlVec *claims_out, Fl *blind_post) {
  FlVec blinds_poly = tape.random_vector("blinds_poly", num_rounds);
  FlVec blinds_evals = tape.random_vector("blinds_evals", num_rounds);
  Fl claim_per_round = claim;
  Comp comm_claim_per_round = compress(commit_scalar(claim_per_round, blind_claim, gens_1));
  ZKSumcheckInstanceProof out;
  FlVec r;
  for (size_t j = 0; j < num_rounds; j++) {
    Fl e0 = fl_zero(), e2 = fl_zero();
    size_t len = A.len / 2;
    for (size_t i = 0; i < len; i++) {
      e0 += A[i] * B[i];
      Fl a2 = A[len + i] + A[len + i] - A[i], b2 = B[len + i] + B[len + i] - B[i];
      e2 += a2 * b2;
    }
    UniPoly poly = UniPoly::from_evals({e0, claim_per_round - e0, e2});
    Comp comm_poly = compress(commit_vec(poly.coeffs.data(), poly.coeffs.size(), blinds_poly[j], gens_n));
    t.append_point("comm_poly", comm_poly.data());
    out.comm_polys.push_back(comm_poly);
    Fl r_j = t.challenge_scalar("challenge_nextround");
    A.bound_poly_var_top(r_j);
    B.bound_poly_var_top(r_j);
    zk_round_tail(poly, r_j, j, blind_claim, blinds_poly, blinds_evals, claim_per_round, comm_claim_per_round, gens_1,
                  gens_n, t, tape, out);
    r.push_back(r_j);
  }
  *r_out = r;
  *claims_out = {A[0], B[0]};
  *blind_post = blinds_evals[num_rounds - 1];
  return out;
}
// :84-179
static inline bool zk_sumcheck_verify(const ZKSumcheckInstanceProof &p, const Comp &comm_claim, size_t num_rounds,
                                      size_t degree_bound, const MultiCommitGens &gens_1, const MultiCommitGens &gens_n,
                                      Transcript &t, Comp *comm_out, FlVec *r_out) {
  if (gens_n.n != degree_bound + 1) return false;
  if (p.comm_polys.size() != num_rounds || p.comm_evals.size() != num_rounds || p.proofs.size() != num_rounds) return false;
  FlVec r;
  for (size_t i = 0; i < num_rounds; i++) {
    const Comp &comm_poly = p.comm_polys[i];
    t.append_point("comm_poly", comm_poly.data());
    Fl r_i = t.challenge_scalar("challenge_nextround");
    const Comp &comm_claim_per_round = i == 0 ? comm_claim : p.comm_evals[i - 1];
    const Comp &comm_eval = p.comm_evals[i];
    t.append_point("comm_claim_per_round", comm_claim_per_round.data());
    t.append_point("comm_eval", comm_eval.data());
    FlVec w = t.challenge_vector("combine_two_claims_to_one", 2);
    Pt pts[2];
    if (!pt_decompress(comm_claim_per_round.data(), &pts[0]) || !pt_decompress(comm_eval.data(), &pts[1])) return false;
    Comp comm_target = compress(msm(w.data(), pts, 2));
    FlVec a_sc(degree_bound + 1, fl_one());
    a_sc[0] += fl_one();
    FlVec a_eval(degree_bound + 1, fl_one());
    for (size_t j = 1; j < a_eval.size(); j++) a_eval[j] = a_eval[j - 1] * r_i;
    FlVec a(degree_bound + 1);
    for (size_t k = 0; k < a.size(); k++) a[k] = w[0] * a_sc[k] + w[1] * a_eval[k];
    if (!dotproduct_verify(p.proofs[i], gens_1, gens_n, t, a, comm_poly, comm_target)) return false;
    r.push_back(r_i);
  }
  *comm_out = p.comm_evals[p.comm_evals.size() - 1];
  *r_out = r;
  return true;
}

// :254-424, comb_func = A*B*C (SP/product_tree.rs:283-286). Tables are bound in place.
struct BatchedClaims { FlVec prod_left, prod_right; Fl prod_eq; FlVec dotp_left, dotp_right, dotp_weight; };
static inline SumcheckInstanceProof prove_cubic_batched(const Fl &claim, size_t num_rounds,
                                                        std::vector<DensePoly *> &A_par, std::vector<DensePoly *> &B_par,
                                                        DensePoly &C_par, std::vector<DensePoly *> &A_seq,
                                                        std::vector<DensePoly *> &B_seq, std::vector<DensePoly *> &C_seq,
                                                        const FlVec &coeffs, Transcript &t, FlVec *r_out,
                                                        BatchedClaims *claims) {
  Fl e = claim;
  FlVec r;
  SumcheckInstanceProof proof;
  auto eval3 = [](const DensePoly &A, const DensePoly &B, const DensePoly &C, Fl *o0, Fl *o2, Fl *o3) {
    Fl e0 = fl_zero(), e2 = fl_zero(), e3 = fl_zero();
    size_t len = A.len / 2;
    for (size_t i = 0; i < len; i++) {
      e0 += A[i] * B[i] * C[i];
      Fl a2 = A[len + i] + A[len + i] - A[i], b2 = B[len + i] + B[len + i] - B[i], c2 = C[len + i] + C[len + i] - C[i];
      e2 += a2 * b2 * c2;
      Fl a3 = a2 + A[len + i] - A[i], b3 = b2 + B[len + i] - B[i], c3 = c2 + C[len + i] - C[i];
      e3 += a3 * b3 * c3;
    }
    *o0 = e0; *o2 = e2; *o3 = e3;
  };
  for (size_t j = 0; j < num_rounds; j++) {
    std::vector<std::array<Fl, 3>> evals;
    for (size_t k = 0; k < A_par.size(); k++) {
      std::array<Fl, 3> ev;
      eval3(*A_par[k], *B_par[k], C_par, &ev[0], &ev[1], &ev[2]);
      evals.push_back(ev);
    }
    for (size_t k = 0; k < A_seq.size(); k++) {
      std::array<Fl, 3> ev;
      eval3(*A_seq[k], *B_seq[k], *C_seq[k], &ev[0], &ev[1], &ev[2]);
      evals.push_back(ev);
    }
    Fl c0 = fl_zero(), c2 = fl_zero(), c3 = fl_zero();
    for (size_t i = 0; i < evals.size(); i++) { c0 += evals[i][0] * coeffs[i]; c2 += evals[i][1] * coeffs[i]; c3 += evals[i][2] * coeffs[i]; }
    UniPoly poly = UniPoly::from_evals({c0, e - c0, c2, c3});
    poly.append_to_transcript(t, "poly");
    Fl r_j = t.challenge_scalar("challenge_nextround");
    r.push_back(r_j);
    for (size_t k = 0; k < A_par.size(); k++) { A_par[k]->bound_poly_var_top(r_j); B_par[k]->bound_poly_var_top(r_j); }
    C_par.bound_poly_var_top(r_j);
    for (size_t k = 0; k < A_seq.size(); k++) {
      A_seq[k]->bound_poly_var_top(r_j);
      B_seq[k]->bound_poly_var_top(r_j);
      C_seq[k]->bound_poly_var_top(r_j);
    }
    e = poly.evaluate(r_j);
    proof.compressed_polys.push_back(poly.compress());
  }
  claims->prod_left.clear(); claims->prod_right.clear();
  for (size_t k = 0; k < A_par.size(); k++) { claims->prod_left.push_back((*A_par[k])[0]); claims->prod_right.push_back((*B_par[k])[0]); }
  claims->prod_eq = C_par[0];
  claims->dotp_left.clear(); claims->dotp_right.clear(); claims->dotp_weight.clear();
  for (size_t k = 0; k < A_seq.size(); k++) {
    claims->dotp_left.push_back((*A_seq[k])[0]);
    claims->dotp_right.push_back((*B_seq[k])[0]);
    claims->dotp_weight.push_back((*C_seq[k])[0]);
  }
  *r_out = r;
  return proof;
}

// ----------------------------------------------------------------------------------------------
// SP/product_tree.rs
struct ProductCircuit {
  std::vector<DensePoly> left_vec, right_vec;
  // :36-56 with compute_layer :18-34
  explicit ProductCircuit(const DensePoly &poly) {
    size_t num_layers = log_2(poly.len);
    size_t half = poly.len / 2;
    left_vec.emplace_back(FlVec(poly.Z.begin(), poly.Z.begin() + half));
    right_vec.emplace_back(FlVec(poly.Z.begin() + half, poly.Z.begin() + 2 * half));
    for (size_t i = 0; i + 1 < num_layers; i++) {
      const DensePoly &l = left_vec[i], &r = right_vec[i];
      size_t len = l.len + r.len;
      FlVec ol(len / 4), orr(len / 4);
      for (size_t k = 0; k < len / 4; k++) ol[k] = l[k] * r[k];
      for (size_t k = len / 4; k < len / 2; k++) orr[k - len / 4] = l[k] * r[k];
      left_vec.emplace_back(std::move(ol));
      right_vec.emplace_back(std::move(orr));
    }
  }
  // :58-63
  Fl evaluate() const {
    size_t len = left_vec.size();
    assert(left_vec[len - 1].num_vars == 0 && right_vec[len - 1].num_vars == 0);
    return left_vec[len - 1][0] * right_vec[len - 1][0];
  }
};
// :66-108
struct DotProductCircuit {
  DensePoly left, right, weight;
  Fl evaluate() const {
    Fl s = fl_zero();
    for (size_t i = 0; i < left.len; i++) s += left[i] * right[i] * weight[i];
    return s;
  }
};
static inline void dense_split(const DensePoly &p, size_t idx, DensePoly *a, DensePoly *b) {
  assert(idx < p.len);
  *a = DensePoly(FlVec(p.Z.begin(), p.Z.begin() + idx));
  *b = DensePoly(FlVec(p.Z.begin() + idx, p.Z.begin() + 2 * idx));
}
static inline void dotp_split(const DotProductCircuit &c, DotProductCircuit *a, DotProductCircuit *b) {
  size_t idx = c.left.len / 2;
  dense_split(c.left, idx, &a->left, &b->left);
  dense_split(c.right, idx, &a->right, &b->right);
  dense_split(c.weight, idx, &a->weight, &b->weight);
}
struct LayerProofBatched { SumcheckInstanceProof proof; FlVec claims_prod_left, claims_prod_right; };
static inline void io(Ar &a, LayerProofBatched &p) { io(a, p.proof); io(a, p.claims_prod_left); io(a, p.claims_prod_right); }
struct ProductCircuitEvalProofBatched { std::vector<LayerProofBatched> proof; FlVec claims_dotp[3]; };
static inline void io(Ar &a, ProductCircuitEvalProofBatched &p) {
  io(a, p.proof); io(a, p.claims_dotp[0]); io(a, p.claims_dotp[1]); io(a, p.claims_dotp[2]);
}
// :259-383
static inline ProductCircuitEvalProofBatched pcepb_prove(std::vector<ProductCircuit *> &prod_circuit_vec,
                                                         std::vector<DotProductCircuit *> &dotp_circuit_vec,
                                                         Transcript &t, FlVec *rand_out) {
  assert(!prod_circuit_vec.empty());
  ProductCircuitEvalProofBatched out;
  size_t num_layers = prod_circuit_vec[0]->left_vec.size();
  FlVec claims_to_verify;
  for (ProductCircuit *c : prod_circuit_vec) claims_to_verify.push_back(c->evaluate());
  FlVec rand;
  for (size_t layer_id = num_layers; layer_id-- > 0;) {
    size_t len = prod_circuit_vec[0]->left_vec[layer_id].len + prod_circuit_vec[0]->right_vec[layer_id].len;
    DensePoly poly_C_par(eq_evals(rand));
    assert(poly_C_par.len == len / 2);
    size_t num_rounds_prod = log_2(poly_C_par.len);
    std::vector<DensePoly *> A_par, B_par, A_seq, B_seq, C_seq;
    for (ProductCircuit *c : prod_circuit_vec) { A_par.push_back(&c->left_vec[layer_id]); B_par.push_back(&c->right_vec[layer_id]); }
    if (layer_id == 0 && !dotp_circuit_vec.empty()) {
      for (DotProductCircuit *d : dotp_circuit_vec) {
        claims_to_verify.push_back(d->evaluate());
        assert(len / 2 == d->left.len && len / 2 == d->right.len && len / 2 == d->weight.len);
      }
      for (DotProductCircuit *d : dotp_circuit_vec) { A_seq.push_back(&d->left); B_seq.push_back(&d->right); C_seq.push_back(&d->weight); }
    }
    FlVec coeff_vec = t.challenge_vector("rand_coeffs_next_layer", claims_to_verify.size());
    Fl claim = fl_zero();
    for (size_t i = 0; i < claims_to_verify.size(); i++) claim += claims_to_verify[i] * coeff_vec[i];
    FlVec rand_prod;
    BatchedClaims cl;
    SumcheckInstanceProof proof =
        prove_cubic_batched(claim, num_rounds_prod, A_par, B_par, poly_C_par, A_seq, B_seq, C_seq, coeff_vec, t, &rand_prod, &cl);
    for (size_t i = 0; i < prod_circuit_vec.size(); i++) {
      t.append_scalar("claim_prod_left", cl.prod_left[i]);
      t.append_scalar("claim_prod_right", cl.prod_right[i]);
    }
    if (layer_id == 0 && !dotp_circuit_vec.empty()) {
      for (size_t i = 0; i < dotp_circuit_vec.size(); i++) {
        t.append_scalar("claim_dotp_left", cl.dotp_left[i]);
        t.append_scalar("claim_dotp_right", cl.dotp_right[i]);
        t.append_scalar("claim_dotp_weight", cl.dotp_weight[i]);
      }
      out.claims_dotp[0] = cl.dotp_left; out.claims_dotp[1] = cl.dotp_right; out.claims_dotp[2] = cl.dotp_weight;
    }
    Fl r_layer = t.challenge_scalar("challenge_r_layer");
    claims_to_verify.clear();
    for (size_t i = 0; i < prod_circuit_vec.size(); i++)
      claims_to_verify.push_back(cl.prod_left[i] + r_layer * (cl.prod_right[i] - cl.prod_left[i]));
    FlVec ext = {r_layer};
    ext.insert(ext.end(), rand_prod.begin(), rand_prod.end());
    rand = ext;
    out.proof.push_back(LayerProofBatched{proof, cl.prod_left, cl.prod_right});
  }
  *rand_out = rand;
  return out;
}
// :385-485
static inline bool pcepb_verify(const ProductCircuitEvalProofBatched &p, const FlVec &claims_prod_vec,
                                const FlVec &claims_dotp_vec, size_t len, Transcript &t, FlVec *claims_out,
                                FlVec *claims_dotp_out, FlVec *rand_out) {
  size_t num_layers = log_2(len);
  FlVec rand;
  if (p.proof.size() != num_layers) return false;
  FlVec claims_to_verify = claims_prod_vec, claims_to_verify_dotp;
  for (size_t i = 0; i < num_layers; i++) {
    size_t num_rounds = i;
    if (i == num_layers - 1) claims_to_verify.insert(claims_to_verify.end(), claims_dotp_vec.begin(), claims_dotp_vec.end());
    FlVec coeff_vec = t.challenge_vector("rand_coeffs_next_layer", claims_to_verify.size());
    Fl claim = fl_zero();
    for (size_t k = 0; k < claims_to_verify.size(); k++) claim += claims_to_verify[k] * coeff_vec[k];
    Fl claim_last;
    FlVec rand_prod;
    if (!sumcheck_verify(p.proof[i].proof, claim, num_rounds, 3, t, &claim_last, &rand_prod)) return false;
    const FlVec &cpl = p.proof[i].claims_prod_left, &cpr = p.proof[i].claims_prod_right;
    if (cpl.size() != claims_prod_vec.size() || cpr.size() != claims_prod_vec.size()) return false;
    for (size_t k = 0; k < claims_prod_vec.size(); k++) {
      t.append_scalar("claim_prod_left", cpl[k]);
      t.append_scalar("claim_prod_right", cpr[k]);
    }
    if (rand.size() != rand_prod.size()) return false;
    Fl eq = eq_evaluate(rand, rand_prod);
    Fl claim_expected = fl_zero();
    for (size_t k = 0; k < claims_prod_vec.size(); k++) claim_expected += coeff_vec[k] * (cpl[k] * cpr[k] * eq);
    if (i == num_layers - 1) {
      size_t npi = claims_prod_vec.size();
      const FlVec &dl = p.claims_dotp[0], &dr = p.claims_dotp[1], &dw = p.claims_dotp[2];
      if (dl.size() != claims_dotp_vec.size() || dr.size() != dl.size() || dw.size() != dl.size()) return false;
      for (size_t k = 0; k < dl.size(); k++) {
        t.append_scalar("claim_dotp_left", dl[k]);
        t.append_scalar("claim_dotp_right", dr[k]);
        t.append_scalar("claim_dotp_weight", dw[k]);
        claim_expected += coeff_vec[k + npi] * dl[k] * dr[k] * dw[k];
      }
    }
    if (claim_expected != claim_last) return false;
    Fl r_layer = t.challenge_scalar("challenge_r_layer");
    claims_to_verify.clear();
    for (size_t k = 0; k < cpl.size(); k++) claims_to_verify.push_back(cpl[k] + r_layer * (cpr[k] - cpl[k]));
    if (i == num_layers - 1) {
      const FlVec &dl = p.claims_dotp[0], &dr = p.claims_dotp[1], &dw = p.claims_dotp[2];
      for (size_t k = 0; k < claims_dotp_vec.size() / 2; k++) {
        claims_to_verify_dotp.push_back(dl[2 * k] + r_layer * (dl[2 * k + 1] - dl[2 * k]));
        claims_to_verify_dotp.push_back(dr[2 * k] + r_layer * (dr[2 * k + 1] - dr[2 * k]));
        claims_to_verify_dotp.push_back(dw[2 * k] + r_layer * (dw[2 * k + 1] - dw[2 * k]));
      }
    }
    FlVec ext = {r_layer};
    ext.insert(ext.end(), rand_prod.begin(), rand_prod.end());
    rand = ext;
  }
  *claims_out = claims_to_verify;
  *claims_dotp_out = claims_to_verify_dotp;
  *rand_out = rand;
  return true;
}

// ----------------------------------------------------------------------------------------------
// SP/sparse_mlpoly.rs
struct SparseMatEntry { size_t row, col; Fl val; };
struct SparseMatPolynomial {
  size_t num_vars_x, num_vars_y;
  std::vector<SparseMatEntry> M;
  size_t get_num_nz_entries() const { return next_pow2(M.size()); }  // :364-366
  // :467-481
  FlVec multiply_vec(size_t num_rows, size_t num_cols, const FlVec &z) const {
    assert(z.size() == num_cols);
    FlVec Mz(num_rows, fl_zero());
    for (const SparseMatEntry &e : M) Mz[e.row] += e.val * z[e.col];
    return Mz;
  }
  // :483-498
  FlVec compute_eval_table_sparse(const FlVec &rx, size_t num_rows, size_t num_cols) const {
    assert(rx.size() == num_rows);
    FlVec out(num_cols, fl_zero());
    for (const SparseMatEntry &e : M) out[e.col] += rx[e.row] * e.val;
    return out;
  }
  // :440-452
  Fl evaluate_with_tables(const FlVec &trx, const FlVec &try_) const {
    assert(pow2(num_vars_x) == trx.size() && pow2(num_vars_y) == try_.size());
    Fl s = fl_zero();
    for (const SparseMatEntry &e : M) s += trx[e.row] * try_[e.col] * e.val;
    return s;
  }
};

// :224-283
struct AddrTimestamps {
  std::vector<std::vector<size_t>> ops_addr_usize;
  std::vector<DensePoly> ops_addr, read_ts;
  DensePoly audit_ts;
  AddrTimestamps() {}
  AddrTimestamps(size_t num_cells, size_t num_ops, std::vector<std::vector<size_t>> ops) {
    std::vector<size_t> audit(num_cells, 0);
    for (auto &inst : ops) {
      assert(inst.size() == num_ops);
      std::vector<size_t> rts(num_ops, 0);
      for (size_t i = 0; i < num_ops; i++) {
        size_t addr = inst[i];
        assert(addr < num_cells);
        size_t r_ts = audit[addr];
        rts[i] = r_ts;
        audit[addr] = r_ts + 1;
      }
      ops_addr.push_back(dense_from_usize(inst));
      read_ts.push_back(dense_from_usize(rts));
    }
    ops_addr_usize = std::move(ops);
    audit_ts = dense_from_usize(audit);
  }
  // :267-282
  std::vector<DensePoly> deref(const FlVec &mem_val) const {
    std::vector<DensePoly> out;
    for (auto &addr : ops_addr_usize) {
      FlVec v(addr.size());
      for (size_t i = 0; i < addr.size(); i++) v[i] = mem_val[addr[i]];
      out.emplace_back(std::move(v));
    }
    return out;
  }
};
// :285-292, :382-438
struct MultiSparseMatPolynomialAsDense {
  size_t batch_size;
  std::vector<DensePoly> val;
  AddrTimestamps row, col;
  DensePoly comb_ops, comb_mem;
};
static inline MultiSparseMatPolynomialAsDense multi_sparse_to_dense_rep(const std::vector<const SparseMatPolynomial *> &polys) {
  assert(!polys.empty());
  size_t N = 0;
  for (auto *p : polys) N = p->get_num_nz_entries() > N ? p->get_num_nz_entries() : N;
  std::vector<std::vector<size_t>> ops_row_vec, ops_col_vec;
  MultiSparseMatPolynomialAsDense d;
  for (auto *p : polys) {
    std::vector<size_t> ops_row(N, 0), ops_col(N, 0);  // :368-380
    FlVec val(N, fl_zero());
    for (size_t i = 0; i < p->M.size(); i++) { ops_row[i] = p->M[i].row; ops_col[i] = p->M[i].col; val[i] = p->M[i].val; }
    ops_row_vec.push_back(std::move(ops_row));
    ops_col_vec.push_back(std::move(ops_col));
    d.val.emplace_back(std::move(val));
  }
  const SparseMatPolynomial *any = polys[0];
  size_t num_mem_cells = any->num_vars_x > any->num_vars_y ? pow2(any->num_vars_x) : pow2(any->num_vars_y);
  d.row = AddrTimestamps(num_mem_cells, N, std::move(ops_row_vec));
  d.col = AddrTimestamps(num_mem_cells, N, std::move(ops_col_vec));
  std::vector<const DensePoly *> parts;
  for (auto &p : d.row.ops_addr) parts.push_back(&p);
  for (auto &p : d.row.read_ts) parts.push_back(&p);
  for (auto &p : d.col.ops_addr) parts.push_back(&p);
  for (auto &p : d.col.read_ts) parts.push_back(&p);
  for (auto &p : d.val) parts.push_back(&p);
  d.comb_ops = dense_merge(parts);
  FlVec cm = d.row.audit_ts.Z;  // :427-428 clone + extend
  cm.insert(cm.end(), d.col.audit_ts.Z.begin(), d.col.audit_ts.Z.end());
  d.comb_mem = DensePoly(std::move(cm));
  d.batch_size = polys.size();
  return d;
}
// :330-352
struct SparseMatPolyCommitment { uint64_t batch_size, num_ops, num_mem_cells; PolyCommitment comm_comb_ops, comm_comb_mem; };
static inline void io(Ar &a, SparseMatPolyCommitment &c) {
  io(a, c.batch_size); io(a, c.num_ops); io(a, c.num_mem_cells); io(a, c.comm_comb_ops); io(a, c.comm_comb_mem);
}
// :500-520
static inline SparseMatPolyCommitment multi_commit(const std::vector<const SparseMatPolynomial *> &polys,
                                                   const SparseMatPolyCommitmentGens &gens,
                                                   MultiSparseMatPolynomialAsDense *dense_out) {
  *dense_out = multi_sparse_to_dense_rep(polys);
  SparseMatPolyCommitment c;
  c.batch_size = polys.size();
  c.comm_comb_ops = dense_commit(dense_out->comb_ops, gens.gens_ops, nullptr, nullptr);
  c.comm_comb_mem = dense_commit(dense_out->comb_mem, gens.gens_mem, nullptr, nullptr);
  c.num_mem_cells = dense_out->row.audit_ts.len;
  c.num_ops = dense_out->row.read_ts[0].len;
  return c;
}

// :42-77
struct Derefs { std::vector<DensePoly> row_ops_val, col_ops_val; DensePoly comb; };
static inline Derefs derefs_new(std::vector<DensePoly> row, std::vector<DensePoly> col) {
  Derefs d;
  d.row_ops_val = std::move(row);
  d.col_ops_val = std::move(col);
  std::vector<const DensePoly *> parts;
  for (auto &p : d.row_ops_val) parts.push_back(&p);
  for (auto &p : d.col_ops_val) parts.push_back(&p);
  d.comb = dense_merge(parts);
  return d;
}
struct DerefsCommitment { PolyCommitment comm_ops_val; };
static inline void io(Ar &a, DerefsCommitment &c) { io(a, c.comm_ops_val); }
// :216-222
static inline void append_derefs_commitment(Transcript &t, const char *label, const DerefsCommitment &c) {
  t.append_message("derefs_commitment", "begin_derefs_commitment");
  append_poly_commitment(t, label, c.comm_ops_val);
  t.append_message("derefs_commitment", "end_derefs_commitment");
}
struct DerefsEvalProof { PolyEvalProof proof_derefs; };
static inline void io(Ar &a, DerefsEvalProof &p) { io(a, p.proof_derefs); }
// :90-158
static inline DerefsEvalProof derefs_eval_prove(const Derefs &derefs, const FlVec &eval_row, const FlVec &eval_col,
                                                const FlVec &r, const PolyCommitmentGens &gens, Transcript &t,
                                                RandomTape &tape) {
  t.append_protocol_name("Derefs evaluation proof");
  FlVec evals = eval_row;
  evals.insert(evals.end(), eval_col.begin(), eval_col.end());
  evals.resize(next_pow2(evals.size()), fl_zero());
  assert(derefs.comb.num_vars == r.size() + log_2(evals.size()));
  t.append_scalars("evals_ops_val", evals);
  FlVec challenges = t.challenge_vector("challenge_combine_n_to_one", log_2(evals.size()));
  DensePoly poly_evals(evals);
  for (size_t i = challenges.size(); i-- > 0;) poly_evals.bound_poly_var_bot(challenges[i]);
  assert(poly_evals.len == 1);
  Fl joint_claim_eval = poly_evals[0];
  FlVec r_joint = challenges;
  r_joint.insert(r_joint.end(), r.begin(), r.end());
  t.append_scalar("joint_claim_eval", joint_claim_eval);
  DerefsEvalProof p;
  p.proof_derefs = polyeval_prove(derefs.comb, nullptr, r_joint, joint_claim_eval, nullptr, gens, t, tape, nullptr);
  return p;
}
// :160-213
static inline bool derefs_eval_verify(const DerefsEvalProof &p, const FlVec &r, const FlVec &eval_row,
                                      const FlVec &eval_col, const PolyCommitmentGens &gens, const DerefsCommitment &comm,
                                      Transcript &t) {
  t.append_protocol_name("Derefs evaluation proof");
  FlVec evals = eval_row;
  evals.insert(evals.end(), eval_col.begin(), eval_col.end());
  evals.resize(next_pow2(evals.size()), fl_zero());
  t.append_scalars("evals_ops_val", evals);
  FlVec challenges = t.challenge_vector("challenge_combine_n_to_one", log_2(evals.size()));
  DensePoly poly_evals(evals);
  for (size_t i = challenges.size(); i-- > 0;) poly_evals.bound_poly_var_bot(challenges[i]);
  Fl joint_claim_eval = poly_evals[0];
  FlVec r_joint = challenges;
  r_joint.insert(r_joint.end(), r.begin(), r.end());
  t.append_scalar("joint_claim_eval", joint_claim_eval);
  return polyeval_verify_plain(p.proof_derefs, gens, t, r_joint, joint_claim_eval, comm.comm_ops_val);
}

// :533-672
struct ProductLayer {
  ProductCircuit *init, *audit;
  std::vector<ProductCircuit *> read_vec, write_vec;
  ~ProductLayer() {
    delete init; delete audit;
    for (auto *c : read_vec) delete c;
    for (auto *c : write_vec) delete c;
  }
  ProductLayer() : init(nullptr), audit(nullptr) {}
  ProductLayer(const ProductLayer &) = delete;
};
// :547-622
static inline void build_hash_layer(const FlVec &eval_table, const std::vector<DensePoly> &addrs_vec,
                                    const std::vector<DensePoly> &derefs_vec, const std::vector<DensePoly> &read_ts_vec,
                                    const DensePoly &audit_ts, const Fl &r_hash, const Fl &r_multiset_check,
                                    DensePoly *init, std::vector<DensePoly> *read, std::vector<DensePoly> *write,
                                    DensePoly *audit) {
  Fl r_hash_sqr = r_hash * r_hash;
  auto hash_func = [&](const Fl &addr, const Fl &val, const Fl &ts) { return ts * r_hash_sqr + val * r_hash + addr; };
  size_t num_mem_cells = eval_table.size();
  FlVec vi(num_mem_cells), va(num_mem_cells);
  for (size_t i = 0; i < num_mem_cells; i++) {
    Fl ai = fl_from_u64((uint64_t)i);
    vi[i] = hash_func(ai, eval_table[i], fl_zero()) - r_multiset_check;
    va[i] = hash_func(ai, eval_table[i], audit_ts[i]) - r_multiset_check;
  }
  *init = DensePoly(std::move(vi));
  *audit = DensePoly(std::move(va));
  for (size_t k = 0; k < addrs_vec.size(); k++) {
    const DensePoly &addrs = addrs_vec[k], &derefs = derefs_vec[k], &read_ts = read_ts_vec[k];
    assert(addrs.len == derefs.len && addrs.len == read_ts.len);
    size_t num_ops = addrs.len;
    FlVec vr(num_ops), vw(num_ops);
    for (size_t i = 0; i < num_ops; i++) {
      vr[i] = hash_func(addrs[i], derefs[i], read_ts[i]) - r_multiset_check;
      vw[i] = hash_func(addrs[i], derefs[i], read_ts[i] + fl_one()) - r_multiset_check;
    }
    read->emplace_back(std::move(vr));
    write->emplace_back(std::move(vw));
  }
}
// :624-671
static inline void layers_new(const FlVec &eval_table, const AddrTimestamps &at, const std::vector<DensePoly> &poly_ops_val,
                              const Fl &r_hash, const Fl &r_multiset_check, ProductLayer *pl) {
  DensePoly init, audit;
  std::vector<DensePoly> read, write;
  build_hash_layer(eval_table, at.ops_addr, poly_ops_val, at.read_ts, at.audit_ts, r_hash, r_multiset_check, &init,
                   &read, &write, &audit);
  pl->init = new ProductCircuit(init);
  for (auto &p : read) pl->read_vec.push_back(new ProductCircuit(p));
  for (auto &p : write) pl->write_vec.push_back(new ProductCircuit(p));
  pl->audit = new ProductCircuit(audit);
}

// :698-707
struct HashLayerProof {
  FlVec eval_row_addr, eval_row_read_ts; Fl eval_row_audit_ts;
  FlVec eval_col_addr, eval_col_read_ts; Fl eval_col_audit_ts;
  FlVec eval_val;
  FlVec eval_derefs_row, eval_derefs_col;
  PolyEvalProof proof_ops, proof_mem;
  DerefsEvalProof proof_derefs;
};
static inline void io(Ar &a, HashLayerProof &p) {
  io(a, p.eval_row_addr); io(a, p.eval_row_read_ts); io(a, p.eval_row_audit_ts);
  io(a, p.eval_col_addr); io(a, p.eval_col_read_ts); io(a, p.eval_col_audit_ts);
  io(a, p.eval_val);
  io(a, p.eval_derefs_row); io(a, p.eval_derefs_col);
  io(a, p.proof_ops); io(a, p.proof_mem); io(a, p.proof_derefs);
}
// :714-738
static inline void hash_prove_helper(const FlVec &rand_mem, const FlVec &rand_ops, const AddrTimestamps &at, FlVec *addr,
                                     FlVec *rts, Fl *audit) {
  for (auto &p : at.ops_addr) addr->push_back(dense_evaluate(p, rand_ops));
  for (auto &p : at.read_ts) rts->push_back(dense_evaluate(p, rand_ops));
  *audit = dense_evaluate(at.audit_ts, rand_mem);
}
// :740-849
static inline HashLayerProof hash_layer_prove(const FlVec &rand_mem, const FlVec &rand_ops,
                                              const MultiSparseMatPolynomialAsDense &dense, const Derefs &derefs,
                                              const SparseMatPolyCommitmentGens &gens, Transcript &t, RandomTape &tape) {
  t.append_protocol_name("Sparse polynomial hash layer proof");
  HashLayerProof p;
  for (auto &d : derefs.row_ops_val) p.eval_derefs_row.push_back(dense_evaluate(d, rand_ops));
  for (auto &d : derefs.col_ops_val) p.eval_derefs_col.push_back(dense_evaluate(d, rand_ops));
  p.proof_derefs = derefs_eval_prove(derefs, p.eval_derefs_row, p.eval_derefs_col, rand_ops, gens.gens_derefs, t, tape);
  hash_prove_helper(rand_mem, rand_ops, dense.row, &p.eval_row_addr, &p.eval_row_read_ts, &p.eval_row_audit_ts);
  hash_prove_helper(rand_mem, rand_ops, dense.col, &p.eval_col_addr, &p.eval_col_read_ts, &p.eval_col_audit_ts);
  for (auto &v : dense.val) p.eval_val.push_back(dense_evaluate(v, rand_ops));
  FlVec evals_ops;
  evals_ops.insert(evals_ops.end(), p.eval_row_addr.begin(), p.eval_row_addr.end());
  evals_ops.insert(evals_ops.end(), p.eval_row_read_ts.begin(), p.eval_row_read_ts.end());
  evals_ops.insert(evals_ops.end(), p.eval_col_addr.begin(), p.eval_col_addr.end());
  evals_ops.insert(evals_ops.end(), p.eval_col_read_ts.begin(), p.eval_col_read_ts.end());
  evals_ops.insert(evals_ops.end(), p.eval_val.begin(), p.eval_val.end());
  evals_ops.resize(next_pow2(evals_ops.size()), fl_zero());
  t.append_scalars("claim_evals_ops", evals_ops);
  FlVec challenges_ops = t.challenge_vector("challenge_combine_n_to_one", log_2(evals_ops.size()));
  DensePoly poly_evals_ops(evals_ops);
  for (size_t i = challenges_ops.size(); i-- > 0;) poly_evals_ops.bound_poly_var_bot(challenges_ops[i]);
  assert(poly_evals_ops.len == 1);
  Fl joint_claim_eval_ops = poly_evals_ops[0];
  FlVec r_joint_ops = challenges_ops;
  r_joint_ops.insert(r_joint_ops.end(), rand_ops.begin(), rand_ops.end());
  t.append_scalar("joint_claim_eval_ops", joint_claim_eval_ops);
  p.proof_ops = polyeval_prove(dense.comb_ops, nullptr, r_joint_ops, joint_claim_eval_ops, nullptr, gens.gens_ops, t, tape, nullptr);
  FlVec evals_mem = {p.eval_row_audit_ts, p.eval_col_audit_ts};
  t.append_scalars("claim_evals_mem", evals_mem);
  FlVec challenges_mem = t.challenge_vector("challenge_combine_two_to_one", log_2(evals_mem.size()));
  DensePoly poly_evals_mem(evals_mem);
  for (size_t i = challenges_mem.size(); i-- > 0;) poly_evals_mem.bound_poly_var_bot(challenges_mem[i]);
  assert(poly_evals_mem.len == 1);
  Fl joint_claim_eval_mem = poly_evals_mem[0];
  FlVec r_joint_mem = challenges_mem;
  r_joint_mem.insert(r_joint_mem.end(), rand_mem.begin(), rand_mem.end());
  t.append_scalar("joint_claim_eval_mem", joint_claim_eval_mem);
  p.proof_mem = polyeval_prove(dense.comb_mem, nullptr, r_joint_mem, joint_claim_eval_mem, nullptr, gens.gens_mem, t, tape, nullptr);
  return p;
}
// :851-900
static inline bool hash_verify_helper(const FlVec &rand_mem, const Fl &claim_init, const FlVec &claim_read,
                                      const FlVec &claim_write, const Fl &claim_audit, const FlVec &eval_ops_val,
                                      const FlVec &eval_ops_addr, const FlVec &eval_read_ts, const Fl &eval_audit_ts,
                                      const FlVec &r, const Fl &r_hash, const Fl &r_multiset_check) {
  Fl r_hash_sqr = r_hash * r_hash;
  auto hash_func = [&](const Fl &addr, const Fl &val, const Fl &ts) { return ts * r_hash_sqr + val * r_hash + addr; };
  Fl eval_init_addr = identity_poly_evaluate(rand_mem);
  Fl eval_init_val = eq_evaluate(r, rand_mem);
  if (hash_func(eval_init_addr, eval_init_val, fl_zero()) - r_multiset_check != claim_init) return false;
  if (claim_read.size() != eval_ops_addr.size() || claim_write.size() != eval_ops_addr.size() ||
      eval_ops_val.size() != eval_ops_addr.size() || eval_read_ts.size() != eval_ops_addr.size())
    return false;
  for (size_t i = 0; i < eval_ops_addr.size(); i++)
    if (hash_func(eval_ops_addr[i], eval_ops_val[i], eval_read_ts[i]) - r_multiset_check != claim_read[i]) return false;
  for (size_t i = 0; i < eval_ops_addr.size(); i++)
    if (hash_func(eval_ops_addr[i], eval_ops_val[i], eval_read_ts[i] + fl_one()) - r_multiset_check != claim_write[i]) return false;
  if (hash_func(eval_init_addr, eval_init_val, eval_audit_ts) - r_multiset_check != claim_audit) return false;
  return true;
}
struct MemClaims { Fl init; FlVec read, write; Fl audit; };
// :902-1032
static inline bool hash_layer_verify(const HashLayerProof &p, const FlVec &rand_mem, const FlVec &rand_ops,
                                     const MemClaims &claims_row, const MemClaims &claims_col, const FlVec &claims_dotp,
                                     const SparseMatPolyCommitment &comm, const SparseMatPolyCommitmentGens &gens,
                                     const DerefsCommitment &comm_derefs, const FlVec &rx, const FlVec &ry,
                                     const Fl &r_hash, const Fl &r_multiset_check, Transcript &t) {
  t.append_protocol_name("Sparse polynomial hash layer proof");
  if (p.eval_derefs_row.size() != p.eval_derefs_col.size()) return false;
  if (!derefs_eval_verify(p.proof_derefs, rand_ops, p.eval_derefs_row, p.eval_derefs_col, gens.gens_derefs, comm_derefs, t)) return false;
  if (claims_dotp.size() != 3 * p.eval_derefs_row.size() || p.eval_val.size() != p.eval_derefs_row.size()) return false;
  for (size_t i = 0; i < claims_dotp.size() / 3; i++) {
    if (claims_dotp[3 * i] != p.eval_derefs_row[i]) return false;
    if (claims_dotp[3 * i + 1] != p.eval_derefs_col[i]) return false;
    if (claims_dotp[3 * i + 2] != p.eval_val[i]) return false;
  }
  FlVec evals_ops;
  evals_ops.insert(evals_ops.end(), p.eval_row_addr.begin(), p.eval_row_addr.end());
  evals_ops.insert(evals_ops.end(), p.eval_row_read_ts.begin(), p.eval_row_read_ts.end());
  evals_ops.insert(evals_ops.end(), p.eval_col_addr.begin(), p.eval_col_addr.end());
  evals_ops.insert(evals_ops.end(), p.eval_col_read_ts.begin(), p.eval_col_read_ts.end());
  evals_ops.insert(evals_ops.end(), p.eval_val.begin(), p.eval_val.end());
  evals_ops.resize(next_pow2(evals_ops.size()), fl_zero());
  t.append_scalars("claim_evals_ops", evals_ops);
  FlVec challenges_ops = t.challenge_vector("challenge_combine_n_to_one", log_2(evals_ops.size()));
  DensePoly poly_evals_ops(evals_ops);
  for (size_t i = challenges_ops.size(); i-- > 0;) poly_evals_ops.bound_poly_var_bot(challenges_ops[i]);
  Fl joint_claim_eval_ops = poly_evals_ops[0];
  FlVec r_joint_ops = challenges_ops;
  r_joint_ops.insert(r_joint_ops.end(), rand_ops.begin(), rand_ops.end());
  t.append_scalar("joint_claim_eval_ops", joint_claim_eval_ops);
  if (!polyeval_verify_plain(p.proof_ops, gens.gens_ops, t, r_joint_ops, joint_claim_eval_ops, comm.comm_comb_ops)) return false;
  FlVec evals_mem = {p.eval_row_audit_ts, p.eval_col_audit_ts};
  t.append_scalars("claim_evals_mem", evals_mem);
  FlVec challenges_mem = t.challenge_vector("challenge_combine_two_to_one", log_2(evals_mem.size()));
  DensePoly poly_evals_mem(evals_mem);
  for (size_t i = challenges_mem.size(); i-- > 0;) poly_evals_mem.bound_poly_var_bot(challenges_mem[i]);
  Fl joint_claim_eval_mem = poly_evals_mem[0];
  FlVec r_joint_mem = challenges_mem;
  r_joint_mem.insert(r_joint_mem.end(), rand_mem.begin(), rand_mem.end());
  t.append_scalar("joint_claim_eval_mem", joint_claim_eval_mem);
  if (!polyeval_verify_plain(p.proof_mem, gens.gens_mem, t, r_joint_mem, joint_claim_eval_mem, comm.comm_comb_mem)) return false;
  if (!hash_verify_helper(rand_mem, claims_row.init, claims_row.read, claims_row.write, claims_row.audit, p.eval_derefs_row,
                          p.eval_row_addr, p.eval_row_read_ts, p.eval_row_audit_ts, rx, r_hash, r_multiset_check))
    return false;
  if (!hash_verify_helper(rand_mem, claims_col.init, claims_col.read, claims_col.write, claims_col.audit, p.eval_derefs_col,
                          p.eval_col_addr, p.eval_col_read_ts, p.eval_col_audit_ts, ry, r_hash, r_multiset_check))
    return false;
  return true;
}

// :1035-1042
struct ProductLayerProof {
  MemClaims eval_row, eval_col;
  FlVec eval_val_left, eval_val_right;
  ProductCircuitEvalProofBatched proof_mem, proof_ops;
};
static inline void io(Ar &a, MemClaims &m) { io(a, m.init); io(a, m.read); io(a, m.write); io(a, m.audit); }
static inline void io(Ar &a, ProductLayerProof &p) {
  io(a, p.eval_row); io(a, p.eval_col); io(a, p.eval_val_left); io(a, p.eval_val_right); io(a, p.proof_mem); io(a, p.proof_ops);
}
// :1049-1227
static inline ProductLayerProof product_layer_prove(ProductLayer &row, ProductLayer &col,
                                                    const MultiSparseMatPolynomialAsDense &dense, const Derefs &derefs,
                                                    const FlVec &eval, Transcript &t, FlVec *rand_mem, FlVec *rand_ops) {
  t.append_protocol_name("Sparse polynomial product layer proof");
  ProductLayerProof p;
  auto claims_of = [](ProductLayer &l, MemClaims *m) {
    m->init = l.init->evaluate();
    m->audit = l.audit->evaluate();
    for (auto *c : l.read_vec) m->read.push_back(c->evaluate());
    for (auto *c : l.write_vec) m->write.push_back(c->evaluate());
    Fl ws = fl_one(), rs = fl_one();
    for (auto &x : m->write) ws *= x;
    for (auto &x : m->read) rs *= x;
    if (m->init * ws != rs * m->audit) throw std::runtime_error("memory check failed");
  };
  claims_of(row, &p.eval_row);
  t.append_scalar("claim_row_eval_init", p.eval_row.init);
  t.append_scalars("claim_row_eval_read", p.eval_row.read);
  t.append_scalars("claim_row_eval_write", p.eval_row.write);
  t.append_scalar("claim_row_eval_audit", p.eval_row.audit);
  claims_of(col, &p.eval_col);
  t.append_scalar("claim_col_eval_init", p.eval_col.init);
  t.append_scalars("claim_col_eval_read", p.eval_col.read);
  t.append_scalars("claim_col_eval_write", p.eval_col.write);
  t.append_scalar("claim_col_eval_audit", p.eval_col.audit);
  assert(eval.size() == derefs.row_ops_val.size() && eval.size() == dense.val.size());
  std::vector<DotProductCircuit> left_vec(eval.size()), right_vec(eval.size());
  for (size_t i = 0; i < derefs.row_ops_val.size(); i++) {
    DotProductCircuit c{derefs.row_ops_val[i], derefs.col_ops_val[i], dense.val[i]};
    dotp_split(c, &left_vec[i], &right_vec[i]);
    Fl el = left_vec[i].evaluate(), er = right_vec[i].evaluate();
    t.append_scalar("claim_eval_dotp_left", el);
    t.append_scalar("claim_eval_dotp_right", er);
    if (el + er != eval[i]) throw std::runtime_error("dotp eval mismatch");
    p.eval_val_left.push_back(el);
    p.eval_val_right.push_back(er);
  }
  assert(row.read_vec.size() == 3);
  std::vector<ProductCircuit *> ops = {row.read_vec[0], row.read_vec[1], row.read_vec[2], row.write_vec[0],
                                       row.write_vec[1], row.write_vec[2], col.read_vec[0], col.read_vec[1],
                                       col.read_vec[2], col.write_vec[0], col.write_vec[1], col.write_vec[2]};
  std::vector<DotProductCircuit *> dotp = {&left_vec[0], &right_vec[0], &left_vec[1], &right_vec[1], &left_vec[2], &right_vec[2]};
  p.proof_ops = pcepb_prove(ops, dotp, t, rand_ops);
  std::vector<ProductCircuit *> mem = {row.init, row.audit, col.init, col.audit};
  std::vector<DotProductCircuit *> none;
  p.proof_mem = pcepb_prove(mem, none, t, rand_mem);
  return p;
}
// :1229-1322
static inline bool product_layer_verify(const ProductLayerProof &p, size_t num_ops, size_t num_cells, const FlVec &eval,
                                        Transcript &t, FlVec *claims_mem, FlVec *rand_mem, FlVec *claims_ops,
                                        FlVec *claims_dotp, FlVec *rand_ops) {
  t.append_protocol_name("Sparse polynomial product layer proof");
  size_t num_instances = eval.size();
  auto subset = [&](const MemClaims &m) {
    if (m.write.size() != num_instances || m.read.size() != num_instances) return false;
    Fl ws = fl_one(), rs = fl_one();
    for (auto &x : m.write) ws *= x;
    for (auto &x : m.read) rs *= x;
    return m.init * ws == rs * m.audit;
  };
  if (!subset(p.eval_row)) return false;
  t.append_scalar("claim_row_eval_init", p.eval_row.init);
  t.append_scalars("claim_row_eval_read", p.eval_row.read);
  t.append_scalars("claim_row_eval_write", p.eval_row.write);
  t.append_scalar("claim_row_eval_audit", p.eval_row.audit);
  if (!subset(p.eval_col)) return false;
  t.append_scalar("claim_col_eval_init", p.eval_col.init);
  t.append_scalars("claim_col_eval_read", p.eval_col.read);
  t.append_scalars("claim_col_eval_write", p.eval_col.write);
  t.append_scalar("claim_col_eval_audit", p.eval_col.audit);
  if (p.eval_val_left.size() != num_instances || p.eval_val_right.size() != num_instances) return false;
  FlVec claims_dotp_circuit;
  for (size_t i = 0; i < num_instances; i++) {
    if (p.eval_val_left[i] + p.eval_val_right[i] != eval[i]) return false;
    t.append_scalar("claim_eval_dotp_left", p.eval_val_left[i]);
    t.append_scalar("claim_eval_dotp_right", p.eval_val_right[i]);
    claims_dotp_circuit.push_back(p.eval_val_left[i]);
    claims_dotp_circuit.push_back(p.eval_val_right[i]);
  }
  FlVec claims_prod_circuit;
  claims_prod_circuit.insert(claims_prod_circuit.end(), p.eval_row.read.begin(), p.eval_row.read.end());
  claims_prod_circuit.insert(claims_prod_circuit.end(), p.eval_row.write.begin(), p.eval_row.write.end());
  claims_prod_circuit.insert(claims_prod_circuit.end(), p.eval_col.read.begin(), p.eval_col.read.end());
  claims_prod_circuit.insert(claims_prod_circuit.end(), p.eval_col.write.begin(), p.eval_col.write.end());
  if (!pcepb_verify(p.proof_ops, claims_prod_circuit, claims_dotp_circuit, num_ops, t, claims_ops, claims_dotp, rand_ops)) return false;
  FlVec unused;
  if (!pcepb_verify(p.proof_mem, {p.eval_row.init, p.eval_row.audit, p.eval_col.init, p.eval_col.audit}, {}, num_cells, t,
                    claims_mem, &unused, rand_mem))
    return false;
  return true;
}

// :1325-1441
struct PolyEvalNetworkProof { ProductLayerProof proof_prod_layer; HashLayerProof proof_hash_layer; };
static inline void io(Ar &a, PolyEvalNetworkProof &p) { io(a, p.proof_prod_layer); io(a, p.proof_hash_layer); }
struct SparseMatPolyEvalProof { DerefsCommitment comm_derefs; PolyEvalNetworkProof poly_eval_network_proof; };
static inline void io(Ar &a, SparseMatPolyEvalProof &p) { io(a, p.comm_derefs); io(a, p.poly_eval_network_proof); }

// :1448-1464
static inline void equalize(const FlVec &rx, const FlVec &ry, FlVec *rx_ext, FlVec *ry_ext) {
  *rx_ext = rx; *ry_ext = ry;
  if (rx.size() < ry.size()) { rx_ext->assign(ry.size() - rx.size(), fl_zero()); rx_ext->insert(rx_ext->end(), rx.begin(), rx.end()); }
  else if (rx.size() > ry.size()) { ry_ext->assign(rx.size() - ry.size(), fl_zero()); ry_ext->insert(ry_ext->end(), ry.begin(), ry.end()); }
}

struct PhaseTimes { double ms[16]; };
static inline double now_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
enum { T_POLYCOMMIT = 0, T_SC1, T_SC2, T_POLYEVAL, T_R1CSPROOF, T_EVAL_SPARSE, T_COMMIT_NONDET, T_BUILD_NET, T_EVALPROOF_NET,
       T_R1CSEVAL, T_PROVE, T_ENCODE, T_NTIMERS };
static const char *const TIMER_NAMES[T_NTIMERS] = {"polycommit", "prove_sc_phase_one", "prove_sc_phase_two", "polyeval",
                                                    "R1CSProof::prove", "eval_sparse_polys", "commit_nondet_witness",
                                                    "build_layered_network", "evalproof_layered_network",
                                                    "R1CSEvalProof::prove", "SNARK::prove", "SNARK::encode"};

// :1466-1533 (and PolyEvalNetworkProof::prove :1336-1370)
static inline SparseMatPolyEvalProof sparse_polyeval_prove(const MultiSparseMatPolynomialAsDense &dense, const FlVec &rx,
                                                           const FlVec &ry, const FlVec &evals,
                                                           const SparseMatPolyCommitmentGens &gens, Transcript &t,
                                                           RandomTape &tape, PhaseTimes *pt) {
  t.append_protocol_name("Sparse polynomial evaluation proof");
  assert(evals.size() == dense.batch_size);
  FlVec rx_ext, ry_ext;
  equalize(rx, ry, &rx_ext, &ry_ext);
  FlVec mem_rx = eq_evals(rx_ext), mem_ry = eq_evals(ry_ext);
  Derefs derefs = derefs_new(dense.row.deref(mem_rx), dense.col.deref(mem_ry));  // :525-530
  double t0 = now_ms();
  SparseMatPolyEvalProof out;
  out.comm_derefs.comm_ops_val = dense_commit(derefs.comb, gens.gens_derefs, nullptr, nullptr);
  append_derefs_commitment(t, "comm_poly_row_col_ops_val", out.comm_derefs);
  if (pt) pt->ms[T_COMMIT_NONDET] = now_ms() - t0;
  FlVec r_mem_check = t.challenge_vector("challenge_r_hash", 2);
  t0 = now_ms();
  ProductLayer row_layer, col_layer;
  layers_new(mem_rx, dense.row, derefs.row_ops_val, r_mem_check[0], r_mem_check[1], &row_layer);
  layers_new(mem_ry, dense.col, derefs.col_ops_val, r_mem_check[0], r_mem_check[1], &col_layer);
  if (pt) pt->ms[T_BUILD_NET] = now_ms() - t0;
  t0 = now_ms();
  t.append_protocol_name("Sparse polynomial evaluation proof");  // PolyEvalNetworkProof::protocol_name :1333,1345
  FlVec rand_mem, rand_ops;
  out.poly_eval_network_proof.proof_prod_layer = product_layer_prove(row_layer, col_layer, dense, derefs, evals, t, &rand_mem, &rand_ops);
  out.poly_eval_network_proof.proof_hash_layer = hash_layer_prove(rand_mem, rand_ops, dense, derefs, gens, t, tape);
  if (pt) pt->ms[T_EVALPROOF_NET] = now_ms() - t0;
  return out;
}
// :1372-1433 and :1535-1571
static inline bool sparse_polyeval_verify(const SparseMatPolyEvalProof &p, const SparseMatPolyCommitment &comm,
                                          const FlVec &rx, const FlVec &ry, const FlVec &evals,
                                          const SparseMatPolyCommitmentGens &gens, Transcript &t) {
  t.append_protocol_name("Sparse polynomial evaluation proof");
  FlVec rx_ext, ry_ext;
  equalize(rx, ry, &rx_ext, &ry_ext);
  size_t nz = comm.num_ops, num_mem_cells = comm.num_mem_cells;
  if (pow2(rx_ext.size()) != num_mem_cells) return false;
  append_derefs_commitment(t, "comm_poly_row_col_ops_val", p.comm_derefs);
  FlVec r_mem_check = t.challenge_vector("challenge_r_hash", 2);
  // PolyEvalNetworkProof::verify
  t.append_protocol_name("Sparse polynomial evaluation proof");
  size_t num_instances = evals.size();
  size_t num_ops = next_pow2(nz), num_cells = pow2(rx_ext.size());
  FlVec claims_mem, rand_mem, claims_ops, claims_dotp, rand_ops;
  if (!product_layer_verify(p.poly_eval_network_proof.proof_prod_layer, num_ops, num_cells, evals, t, &claims_mem, &rand_mem,
                            &claims_ops, &claims_dotp, &rand_ops))
    return false;
  if (claims_mem.size() != 4 || claims_ops.size() != 4 * num_instances || claims_dotp.size() != 3 * num_instances) return false;
  MemClaims cr, cc;
  cr.init = claims_mem[0]; cr.audit = claims_mem[1];
  cr.read.assign(claims_ops.begin(), claims_ops.begin() + num_instances);
  cr.write.assign(claims_ops.begin() + num_instances, claims_ops.begin() + 2 * num_instances);
  cc.init = claims_mem[2]; cc.audit = claims_mem[3];
  cc.read.assign(claims_ops.begin() + 2 * num_instances, claims_ops.begin() + 3 * num_instances);
  cc.write.assign(claims_ops.begin() + 3 * num_instances, claims_ops.begin() + 4 * num_instances);
  return hash_layer_verify(p.poly_eval_network_proof.proof_hash_layer, rand_mem, rand_ops, cr, cc, claims_dotp, comm, gens,
                           p.comm_derefs, rx_ext, ry_ext, r_mem_check[0], r_mem_check[1], t);
}

// ----------------------------------------------------------------------------------------------
// SP/r1csinstance.rs
struct R1CSInstance {
  size_t num_cons, num_vars, num_inputs;
  SparseMatPolynomial A, B, C;
  // :240-270
  bool is_sat(const FlVec &vars, const FlVec &input) const {
    assert(vars.size() == num_vars && input.size() == num_inputs);
    FlVec z = vars;
    z.push_back(fl_one());
    z.insert(z.end(), input.begin(), input.end());
    size_t nc = num_vars + num_inputs + 1;
    FlVec Az = A.multiply_vec(num_cons, nc, z), Bz = B.multiply_vec(num_cons, nc, z), Cz = C.multiply_vec(num_cons, nc, z);
    for (size_t i = 0; i < num_cons; i++)
      if (Az[i] * Bz[i] != Cz[i]) return false;
    return true;
  }
};
struct R1CSCommitment { uint64_t num_cons, num_vars, num_inputs; SparseMatPolyCommitment comm; };
static inline void io(Ar &a, R1CSCommitment &c) { io(a, c.num_cons); io(a, c.num_vars); io(a, c.num_inputs); io(a, c.comm); }

enum R1CSError { R1CS_OK = 0, R1CS_NON_POW2_CONS = 1, R1CS_NON_POW2_VARS = 2, R1CS_INVALID_NUM_INPUTS = 3,
                 R1CS_INVALID_INDEX = 4, R1CS_INVALID_SCALAR = 5 };  // SP/errors.rs:30-45

struct CooEntry { uint64_t row, col; uint8_t val[32]; };
// SP/lib.rs:138-244 (Instance::new; the unused zlib digest :241 is not computed)
static inline R1CSError instance_new(size_t num_cons, size_t num_vars, size_t num_inputs, const CooEntry *A, size_t nA,
                                     const CooEntry *B, size_t nB, const CooEntry *C, size_t nC, R1CSInstance *out) {
  size_t num_vars_padded = num_vars > num_inputs + 1 ? num_vars : num_inputs + 1;
  num_vars_padded = next_pow2(num_vars_padded);
  size_t num_cons_padded = num_cons;
  if (num_cons_padded == 0 || num_cons_padded == 1) num_cons_padded = 2;
  if (next_pow2(num_cons) != num_cons) num_cons_padded = next_pow2(num_cons);
  auto conv = [&](const CooEntry *tups, size_t n, std::vector<SparseMatEntry> *mat) -> R1CSError {
    for (size_t i = 0; i < n; i++) {
      size_t row = tups[i].row, col = tups[i].col;
      if (row >= num_cons) return R1CS_INVALID_INDEX;
      if (col >= num_vars + 1 + num_inputs) return R1CS_INVALID_INDEX;
      Fl val;
      if (!fl_from_bytes(tups[i].val, &val)) return R1CS_INVALID_SCALAR;
      if (col >= num_vars) mat->push_back({row, col + num_vars_padded - num_vars, val});
      else mat->push_back({row, col, val});
    }
    if (num_cons == 0 || num_cons == 1)
      for (size_t i = n; i < num_cons_padded; i++) mat->push_back({i, num_vars, fl_zero()});
    return R1CS_OK;
  };
  R1CSError e;
  out->A.M.clear(); out->B.M.clear(); out->C.M.clear();
  if ((e = conv(A, nA, &out->A.M)) != R1CS_OK) return e;
  if ((e = conv(B, nB, &out->B.M)) != R1CS_OK) return e;
  if ((e = conv(C, nC, &out->C.M)) != R1CS_OK) return e;
  out->num_cons = num_cons_padded;
  out->num_vars = num_vars_padded;
  out->num_inputs = num_inputs;
  size_t nx = log_2(num_cons_padded), ny = log_2(2 * num_vars_padded);
  out->A.num_vars_x = out->B.num_vars_x = out->C.num_vars_x = nx;
  out->A.num_vars_y = out->B.num_vars_y = out->C.num_vars_y = ny;
  return R1CS_OK;
}
// SP/lib.rs:347-358 -> r1csinstance.rs:309-321
static inline R1CSCommitment snark_encode(const R1CSInstance &inst, const SNARKGens &gens,
                                          MultiSparseMatPolynomialAsDense *decomm) {
  R1CSCommitment c;
  c.comm = multi_commit({&inst.A, &inst.B, &inst.C}, gens.gens_r1cs_eval, decomm);
  c.num_cons = inst.num_cons; c.num_vars = inst.num_vars; c.num_inputs = inst.num_inputs;
  return c;
}

// SP/r1csproof.rs:21-47
struct R1CSProof {
  PolyCommitment comm_vars;
  ZKSumcheckInstanceProof sc_proof_phase1;
  Comp claims_phase2[4];
  KnowledgeProof pok_Cz_claim;
  ProductProof proof_prod;
  EqualityProof proof_eq_sc_phase1;
  ZKSumcheckInstanceProof sc_proof_phase2;
  Comp comm_vars_at_ry;
  PolyEvalProof proof_eval_vars_at_ry;
  EqualityProof proof_eq_sc_phase2;
};
static inline void io(Ar &a, R1CSProof &p) {
  io(a, p.comm_vars); io(a, p.sc_proof_phase1);
  for (int i = 0; i < 4; i++) io(a, p.claims_phase2[i]);
  io(a, p.pok_Cz_claim); io(a, p.proof_prod); io(a, p.proof_eq_sc_phase1); io(a, p.sc_proof_phase2);
  io(a, p.comm_vars_at_ry); io(a, p.proof_eval_vars_at_ry); io(a, p.proof_eq_sc_phase2);
}
// SP/lib.rs:330-338 ; R1CSEvalProof { proof } SP/r1csinstance.rs:324-328
struct SNARK { R1CSProof r1cs_sat_proof; Fl inst_evals[3]; SparseMatPolyEvalProof r1cs_eval_proof; };
static inline void io(Ar &a, SNARK &p) {
  io(a, p.r1cs_sat_proof);
  for (int i = 0; i < 3; i++) io(a, p.inst_evals[i]);
  io(a, p.r1cs_eval_proof);
}
template <class T>
static inline std::vector<uint8_t> serialize(T &x) {
  std::vector<uint8_t> out;
  Ar a{true, &out, nullptr, 0, 0, true};
  io(a, x);
  return out;
}
template <class T>
static inline bool deserialize(const uint8_t *b, size_t n, T *x) {
  Ar a{false, nullptr, b, 0, n, true};
  io(a, *x);
  return a.ok && a.pos == n;
}

// VP/commit_test.rs:136-334
static inline R1CSProof my_r1csproof_prove(const R1CSInstance &inst, const FlVec &vars, const FlVec &input,
                                           const R1CSGens &gens, Transcript &t, RandomTape &tape, const DensePoly &poly_vars,
                                           const PolyCommitment &comm_vars, const FlVec &blinds_vars, FlVec *rx_out,
                                           FlVec *ry_out, PhaseTimes *pt) {
  double t_all = now_ms();
  t.append_protocol_name("R1CS proof");
  assert(input.size() < vars.size());
  double t0 = now_ms();
  append_poly_commitment(t, "poly_commitment", comm_vars);
  if (pt) pt->ms[T_POLYCOMMIT] = now_ms() - t0;
  t0 = now_ms();
  size_t num_inputs = input.size(), num_vars = vars.size();
  FlVec z = vars;
  z.push_back(fl_one());
  z.insert(z.end(), input.begin(), input.end());
  z.resize(z.size() + (num_vars - num_inputs - 1), fl_zero());
  size_t num_rounds_x = log_2(inst.num_cons), num_rounds_y = log_2(z.size());  // my_log2 on powers of two
  FlVec tau = t.challenge_vector("challenge_tau", num_rounds_x);
  DensePoly poly_tau(eq_evals(tau));
  DensePoly poly_Az(inst.A.multiply_vec(inst.num_cons, z.size(), z));
  DensePoly poly_Bz(inst.B.multiply_vec(inst.num_cons, z.size(), z));
  DensePoly poly_Cz(inst.C.multiply_vec(inst.num_cons, z.size(), z));
  R1CSProof P;
  P.comm_vars = comm_vars;
  FlVec rx, claims_phase1;
  Fl blind_claim_postsc1;
  // SP/r1csproof.rs:94-127
  P.sc_proof_phase1 = zk_prove_cubic_with_additive_term(fl_zero(), fl_zero(), num_rounds_x, poly_tau, poly_Az, poly_Bz, poly_Cz,
                                                        gens.gens_sc.gens_1, gens.gens_sc.gens_4, t, tape, &rx,
                                                        &claims_phase1, &blind_claim_postsc1);
  assert(poly_tau.len == 1 && poly_Az.len == 1 && poly_Bz.len == 1 && poly_Cz.len == 1);
  if (pt) pt->ms[T_SC1] = now_ms() - t0;
  Fl tau_claim = poly_tau[0], Az_claim = poly_Az[0], Bz_claim = poly_Bz[0], Cz_claim = poly_Cz[0];
  Fl Az_blind = tape.random_scalar("Az_blind"), Bz_blind = tape.random_scalar("Bz_blind"),
     Cz_blind = tape.random_scalar("Cz_blind"), prod_Az_Bz_blind = tape.random_scalar("prod_Az_Bz_blind");
  Comp comm_Cz_claim, comm_Az_claim, comm_Bz_claim, comm_prod_Az_Bz_claims;
  P.pok_Cz_claim = knowledge_prove(gens.gens_sc.gens_1, t, tape, Cz_claim, Cz_blind, &comm_Cz_claim);
  Fl prod = Az_claim * Bz_claim;
  P.proof_prod = product_prove(gens.gens_sc.gens_1, t, tape, Az_claim, Az_blind, Bz_claim, Bz_blind, prod, prod_Az_Bz_blind,
                               &comm_Az_claim, &comm_Bz_claim, &comm_prod_Az_Bz_claims);
  t.append_point("comm_Az_claim", comm_Az_claim.data());
  t.append_point("comm_Bz_claim", comm_Bz_claim.data());
  t.append_point("comm_Cz_claim", comm_Cz_claim.data());
  t.append_point("comm_prod_Az_Bz_claims", comm_prod_Az_Bz_claims.data());
  P.claims_phase2[0] = comm_Az_claim; P.claims_phase2[1] = comm_Bz_claim; P.claims_phase2[2] = comm_Cz_claim;
  P.claims_phase2[3] = comm_prod_Az_Bz_claims;
  Fl taus_bound_rx = tau_claim;
  Fl blind_expected_claim_postsc1 = taus_bound_rx * (prod_Az_Bz_blind - Cz_blind);
  Fl claim_post_phase1 = (Az_claim * Bz_claim - Cz_claim) * taus_bound_rx;
  P.proof_eq_sc_phase1 = equality_prove(gens.gens_sc.gens_1, t, tape, claim_post_phase1, blind_expected_claim_postsc1,
                                        claim_post_phase1, blind_claim_postsc1);
  t0 = now_ms();
  Fl r_A = t.challenge_scalar("challenege_Az"), r_B = t.challenge_scalar("challenege_Bz"), r_C = t.challenge_scalar("challenege_Cz");
  Fl claim_phase2 = r_A * Az_claim + r_B * Bz_claim + r_C * Cz_claim;
  Fl blind_claim_phase2 = r_A * Az_blind + r_B * Bz_blind + r_C * Cz_blind;
  FlVec evals_ABC;
  {
    FlVec evals_rx = eq_evals(rx);
    FlVec eA = inst.A.compute_eval_table_sparse(evals_rx, inst.num_cons, z.size());
    FlVec eB = inst.B.compute_eval_table_sparse(evals_rx, inst.num_cons, z.size());
    FlVec eC = inst.C.compute_eval_table_sparse(evals_rx, inst.num_cons, z.size());
    evals_ABC.resize(eA.size());
    for (size_t i = 0; i < eA.size(); i++) evals_ABC[i] = r_A * eA[i] + r_B * eB[i] + r_C * eC[i];
  }
  FlVec ry, claims_phase2;
  Fl blind_claim_postsc2;
  DensePoly poly_z(z), poly_ABC(evals_ABC);
  // SP/r1csproof.rs:129-155
  P.sc_proof_phase2 = zk_prove_quad(claim_phase2, blind_claim_phase2, num_rounds_y, poly_z, poly_ABC, gens.gens_sc.gens_1,
                                    gens.gens_sc.gens_3, t, tape, &ry, &claims_phase2, &blind_claim_postsc2);
  if (pt) pt->ms[T_SC2] = now_ms() - t0;
  t0 = now_ms();
  FlVec ry1(ry.begin() + 1, ry.end());
  Fl eval_vars_at_ry = dense_evaluate(poly_vars, ry1);
  Fl blind_eval = tape.random_scalar("blind_eval");
  P.proof_eval_vars_at_ry = polyeval_prove(poly_vars, &blinds_vars, ry1, eval_vars_at_ry, &blind_eval, gens.gens_pc, t, tape,
                                           &P.comm_vars_at_ry);
  if (pt) pt->ms[T_POLYEVAL] = now_ms() - t0;
  Fl blind_eval_Z_at_ry = (fl_one() - ry[0]) * blind_eval;
  Fl blind_expected_claim_postsc2 = claims_phase2[1] * blind_eval_Z_at_ry;
  Fl claim_post_phase2 = claims_phase2[0] * claims_phase2[1];
  P.proof_eq_sc_phase2 = equality_prove(gens.gens_pc.gens.gens_1, t, tape, claim_post_phase2, blind_expected_claim_postsc2,
                                        claim_post_phase2, blind_claim_postsc2);
  *rx_out = rx;
  *ry_out = ry;
  if (pt) pt->ms[T_R1CSPROOF] = now_ms() - t_all;
  return P;
}

// VP/commit_test.rs:59-133. tape_seed = the scalar the reference draws from OsRng in RandomTape::new(b"proof") (:74).
static inline SNARK my_lib_prove(const R1CSInstance &inst, const MultiSparseMatPolynomialAsDense &decomm, const FlVec &vars,
                                 const FlVec &inputs, const SNARKGens &gens, Transcript &t, const DensePoly &poly_vars,
                                 const PolyCommitment &comm_vars, const FlVec &blinds_vars, const Fl &tape_seed,
                                 PhaseTimes *pt) {
  double t_all = now_ms();
  RandomTape tape("proof", 5, tape_seed);
  t.append_protocol_name("Spartan SNARK proof");
  SNARK S;
  FlVec rx, ry;
  S.r1cs_sat_proof = my_r1csproof_prove(inst, vars, inputs, gens.gens_r1cs_sat, t, tape, poly_vars, comm_vars, blinds_vars, &rx, &ry, pt);
  double t0 = now_ms();
  {
    // SP/r1csinstance.rs:304-307 -> sparse_mlpoly.rs:454-465
    FlVec trx = eq_evals(rx), try_ = eq_evals(ry);
    S.inst_evals[0] = inst.A.evaluate_with_tables(trx, try_);
    S.inst_evals[1] = inst.B.evaluate_with_tables(trx, try_);
    S.inst_evals[2] = inst.C.evaluate_with_tables(trx, try_);
    t.append_scalar("Ar_claim", S.inst_evals[0]);
    t.append_scalar("Br_claim", S.inst_evals[1]);
    t.append_scalar("Cr_claim", S.inst_evals[2]);
  }
  if (pt) pt->ms[T_EVAL_SPARSE] = now_ms() - t0;
  t0 = now_ms();
  S.r1cs_eval_proof = sparse_polyeval_prove(decomm, rx, ry, {S.inst_evals[0], S.inst_evals[1], S.inst_evals[2]},
                                            gens.gens_r1cs_eval, t, tape, pt);
  if (pt) { pt->ms[T_R1CSEVAL] = now_ms() - t0; pt->ms[T_PROVE] = now_ms() - t_all; }
  return S;
}

// SP/sparse_mlpoly.rs:1586-1622
static inline Fl sparse_poly_evaluate(size_t num_vars, const std::vector<std::pair<size_t, Fl>> &Z, const FlVec &r) {
  assert(num_vars == r.size());
  Fl sum = fl_zero();
  for (auto &e : Z) {
    Fl chi = fl_one();
    for (size_t j = 0; j < r.size(); j++) {
      bool bit = (e.first >> (r.size() - j - 1)) & 1;  // Math::get_bits, MSB first
      chi *= bit ? r[j] : fl_one() - r[j];
    }
    sum += chi * e.second;
  }
  return sum;
}

// VP/commit_test.rs:340-496
static inline bool my_r1csproof_verify(const R1CSProof &P, size_t num_vars, size_t num_cons, const FlVec &input,
                                       const Fl evals[3], Transcript &t, const R1CSGens &gens, const PolyCommitment &comm_1,
                                       const PolyCommitment &comm_2, FlVec *rx_out, FlVec *ry_out) {
  t.append_protocol_name("R1CS proof");
  size_t n = num_vars;
  if (comm_1.C.size() != P.comm_vars.C.size() || comm_2.C.size() != P.comm_vars.C.size()) return false;
  PolyCommitment combine;
  for (size_t i = 0; i < P.comm_vars.C.size(); i++) {
    Pt a, b;
    if (!pt_decompress(comm_1.C[i].data(), &a) || !pt_decompress(comm_2.C[i].data(), &b)) return false;
    combine.C.push_back(compress(pt_add(a, b)));
  }
  append_poly_commitment(t, "poly_commitment", combine);
  size_t num_rounds_x = log_2(num_cons), num_rounds_y = log_2(2 * num_vars);
  FlVec tau = t.challenge_vector("challenge_tau", num_rounds_x);
  Comp claim_phase1 = compress(commit_scalar(fl_zero(), fl_zero(), gens.gens_sc.gens_1));
  Comp comm_claim_post_phase1;
  FlVec rx;
  if (!zk_sumcheck_verify(P.sc_proof_phase1, claim_phase1, num_rounds_x, 3, gens.gens_sc.gens_1, gens.gens_sc.gens_4, t,
                          &comm_claim_post_phase1, &rx))
    return false;
  const Comp &comm_Az_claim = P.claims_phase2[0], &comm_Bz_claim = P.claims_phase2[1], &comm_Cz_claim = P.claims_phase2[2],
             &comm_prod_Az_Bz_claims = P.claims_phase2[3];
  if (!knowledge_verify(P.pok_Cz_claim, gens.gens_sc.gens_1, t, comm_Cz_claim)) return false;
  if (!product_verify(P.proof_prod, gens.gens_sc.gens_1, t, comm_Az_claim, comm_Bz_claim, comm_prod_Az_Bz_claims)) return false;
  t.append_point("comm_Az_claim", comm_Az_claim.data());
  t.append_point("comm_Bz_claim", comm_Bz_claim.data());
  t.append_point("comm_Cz_claim", comm_Cz_claim.data());
  t.append_point("comm_prod_Az_Bz_claims", comm_prod_Az_Bz_claims.data());
  Fl taus_bound_rx = eq_evaluate(rx, tau);
  Pt pAB, pC;
  if (!pt_decompress(comm_prod_Az_Bz_claims.data(), &pAB) || !pt_decompress(comm_Cz_claim.data(), &pC)) return false;
  Comp expected_claim_post_phase1 = compress(pt_mul(taus_bound_rx, pt_sub(pAB, pC)));
  if (!equality_verify(P.proof_eq_sc_phase1, gens.gens_sc.gens_1, t, expected_claim_post_phase1, comm_claim_post_phase1)) return false;
  Fl r_A = t.challenge_scalar("challenege_Az"), r_B = t.challenge_scalar("challenege_Bz"), r_C = t.challenge_scalar("challenege_Cz");
  Pt p3[3];
  if (!pt_decompress(comm_Az_claim.data(), &p3[0]) || !pt_decompress(comm_Bz_claim.data(), &p3[1]) || !pt_decompress(comm_Cz_claim.data(), &p3[2]))
    return false;
  Fl s3[3] = {r_A, r_B, r_C};
  Comp comm_claim_phase2 = compress(msm(s3, p3, 3));
  Comp comm_claim_post_phase2;
  FlVec ry;
  if (!zk_sumcheck_verify(P.sc_proof_phase2, comm_claim_phase2, num_rounds_y, 2, gens.gens_sc.gens_1, gens.gens_sc.gens_3, t,
                          &comm_claim_post_phase2, &ry))
    return false;
  FlVec ry1(ry.begin() + 1, ry.end());
  if (!polyeval_verify(P.proof_eval_vars_at_ry, gens.gens_pc, t, ry1, P.comm_vars_at_ry, P.comm_vars)) return false;
  std::vector<std::pair<size_t, Fl>> entries;
  entries.push_back({0, fl_one()});
  for (size_t i = 0; i < input.size(); i++) entries.push_back({i + 1, input[i]});
  Fl poly_input_eval = sparse_poly_evaluate(log_2(n), entries, ry1);
  Pt cv;
  if (!pt_decompress(P.comm_vars_at_ry.data(), &cv)) return false;
  Fl s2[2] = {fl_one() - ry[0], ry[0]};
  Pt p2[2] = {cv, commit_scalar(poly_input_eval, fl_zero(), gens.gens_pc.gens.gens_1)};
  Pt comm_eval_Z_at_ry = msm(s2, p2, 2);
  Comp expected_claim_post_phase2 = compress(pt_mul(r_A * evals[0] + r_B * evals[1] + r_C * evals[2], comm_eval_Z_at_ry));
  if (!equality_verify(P.proof_eq_sc_phase2, gens.gens_sc.gens_1, t, expected_claim_post_phase2, comm_claim_post_phase2)) return false;
  *rx_out = rx;
  *ry_out = ry;
  return true;
}
// VP/commit_test.rs:498-544
static inline bool my_lib_verify(const SNARK &pf, const R1CSCommitment &comm, const FlVec &input, Transcript &t,
                                 const SNARKGens &gens, const PolyCommitment &com_1, const PolyCommitment &com_2) {
  t.append_protocol_name("Spartan SNARK proof");
  if (input.size() != comm.num_inputs) return false;
  FlVec rx, ry;
  if (!my_r1csproof_verify(pf.r1cs_sat_proof, comm.num_vars, comm.num_cons, input, pf.inst_evals, t, gens.gens_r1cs_sat, com_1,
                           com_2, &rx, &ry))
    return false;
  t.append_scalar("Ar_claim", pf.inst_evals[0]);
  t.append_scalar("Br_claim", pf.inst_evals[1]);
  t.append_scalar("Cr_claim", pf.inst_evals[2]);
  return sparse_polyeval_verify(pf.r1cs_eval_proof, comm.comm, rx, ry, {pf.inst_evals[0], pf.inst_evals[1], pf.inst_evals[2]},
                                gens.gens_r1cs_eval, t);
}

}  // namespace orc
