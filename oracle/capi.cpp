// TEST INFRASTRUCTURE ONLY: C entry points of the CPU oracle for ctypes (tests/, smoke(), bench cpu_baseline).
// Field elements cross as 32-byte little-endian CANONICAL values unless a name says "mont".
#include "vpin.hpp"

using namespace orc;

extern "C" {

// ---- field (KATs of SP/scalar/ristretto255.rs:788-1214 are replayed through these) ----
int orc_fl_from_bytes(const uint8_t *in32, uint8_t *out_mont32) {
  Fl x;
  bool ok = fl_from_bytes(in32, &x);
  memcpy(out_mont32, x.v, 32);
  return ok ? 1 : 0;
}
void orc_fl_to_bytes(const uint8_t *mont32, uint8_t *out32) { Fl x; memcpy(x.v, mont32, 32); fl_to_bytes(x, out32); }
void orc_fl_from_bytes_wide(const uint8_t *in64, uint8_t *out_mont32) { Fl x = fl_from_bytes_wide(in64); memcpy(out_mont32, x.v, 32); }
void orc_fl_binop_mont(int op, const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
  Fl a, b, r;
  memcpy(a.v, a32, 32);
  memcpy(b.v, b32, 32);
  switch (op) {
    case 0: r = fl_add(a, b); break;
    case 1: r = fl_sub(a, b); break;
    case 2: r = fl_mul(a, b); break;
    case 3: r = fl_neg(a); break;
    case 4: r = fl_invert(a); break;
    case 5: r = fl_sqr(a); break;
    default: r = fl_zero();
  }
  memcpy(out32, r.v, 32);
}

// ---- group ----
int orc_pt_decompress_ok(const uint8_t *c32) { Pt p; return pt_decompress(c32, &p) ? 1 : 0; }
int orc_pt_add(const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
  Pt a, b;
  if (!pt_decompress(a32, &a) || !pt_decompress(b32, &b)) return 0;
  pt_compress(pt_add(a, b), out32);
  return 1;
}
void orc_from_uniform_bytes(const uint8_t *in64, uint8_t *out32) { pt_compress(pt_from_uniform_bytes(in64), out32); }
// scalars: n x 32 canonical bytes; points: n x 32 compressed
int orc_msm(uint64_t n, const uint8_t *scalars, const uint8_t *points, uint8_t *out32) {
  std::vector<Fl> s(n);
  std::vector<Pt> p(n);
  for (uint64_t i = 0; i < n; i++) {
    if (!fl_from_bytes(scalars + 32 * i, &s[i])) return 0;
    if (!pt_decompress(points + 32 * i, &p[i])) return 0;
  }
  pt_compress(msm(s.data(), p.data(), n), out32);
  return 1;
}
// SP/commitments.rs:20-38: writes n+1 compressed points (last is h)
void orc_derive_gens(const char *label, uint64_t n, uint8_t *out) {
  MultiCommitGens g = mcg_new(n, label);
  for (uint64_t i = 0; i < n; i++) pt_compress(g.G[i], out + 32 * i);
  pt_compress(g.h, out + 32 * n);
}
void orc_shake256(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t outlen) {
  Shake256 s;
  s.absorb(in, n);
  s.squeeze(out, outlen);
}

// ---- transcript ----
void *orc_transcript_new(const uint8_t *label, uint64_t n) { return new Transcript((const char *)label, n); }
void orc_transcript_free(void *t) { delete (Transcript *)t; }
void orc_transcript_append(void *t, const char *label, const uint8_t *msg, uint64_t n) { ((Transcript *)t)->append_message(label, msg, n); }
void orc_transcript_challenge(void *t, const char *label, uint8_t *out, uint64_t n) { ((Transcript *)t)->challenge_bytes(label, out, n); }
void orc_transcript_challenge_scalar(void *t, const char *label, uint8_t *out32) {
  Fl x = ((Transcript *)t)->challenge_scalar(label);
  fl_to_bytes(x, out32);
}

// ---- polynomial helpers (Montgomery 32-byte elements in and out) ----
static FlVec load_mont(const uint8_t *p, uint64_t n) { FlVec v(n); memcpy((void *)v.data(), p, 32 * n); return v; }
static void store_mont(const FlVec &v, uint8_t *p) { memcpy(p, v.data(), 32 * v.size()); }
void orc_eq_evals_mont(const uint8_t *r, uint64_t ell, uint8_t *out) { store_mont(eq_evals(load_mont(r, ell)), out); }
void orc_bind_top_mont(uint8_t *Z, uint64_t len, const uint8_t *r32) {
  DensePoly p(load_mont(Z, len));
  Fl r;
  memcpy(r.v, r32, 32);
  p.bound_poly_var_top(r);
  memcpy(Z, p.Z.data(), 32 * p.len);
}
// SP/sumcheck.rs:619-652 with comb A*(B*C-D): out = eval_point_0, _2, _3
void orc_cubic_round_mont(const uint8_t *A, const uint8_t *B, const uint8_t *C, const uint8_t *D, uint64_t len, uint8_t *out96) {
  FlVec a = load_mont(A, len), b = load_mont(B, len), c = load_mont(C, len), d = load_mont(D, len);
  Fl e0 = fl_zero(), e2 = fl_zero(), e3 = fl_zero();
  uint64_t h = len / 2;
  for (uint64_t i = 0; i < h; i++) {
    e0 += a[i] * (b[i] * c[i] - d[i]);
    Fl a2 = a[h + i] + a[h + i] - a[i], b2 = b[h + i] + b[h + i] - b[i], c2 = c[h + i] + c[h + i] - c[i], d2 = d[h + i] + d[h + i] - d[i];
    e2 += a2 * (b2 * c2 - d2);
    Fl a3 = a2 + a[h + i] - a[i], b3 = b2 + b[h + i] - b[i], c3 = c2 + c[h + i] - c[i], d3 = d2 + d[h + i] - d[i];
    e3 += a3 * (b3 * c3 - d3);
  }
  memcpy(out96, e0.v, 32); memcpy(out96 + 32, e2.v, 32); memcpy(out96 + 64, e3.v, 32);
}
// SP/sumcheck.rs:456-469 with comb A*B: out = eval_point_0, _2
void orc_quad_round_mont(const uint8_t *A, const uint8_t *B, uint64_t len, uint8_t *out64) {
  FlVec a = load_mont(A, len), b = load_mont(B, len);
  Fl e0 = fl_zero(), e2 = fl_zero();
  uint64_t h = len / 2;
  for (uint64_t i = 0; i < h; i++) {
    e0 += a[i] * b[i];
    e2 += (a[h + i] + a[h + i] - a[i]) * (b[h + i] + b[h + i] - b[i]);
  }
  memcpy(out64, e0.v, 32); memcpy(out64 + 32, e2.v, 32);
}
// SP/sumcheck.rs:296-320 with comb A*B*C: out = eval_point_0, _2, _3
void orc_cubic3_round_mont(const uint8_t *A, const uint8_t *B, const uint8_t *C, uint64_t len, uint8_t *out96) {
  FlVec a = load_mont(A, len), b = load_mont(B, len), c = load_mont(C, len);
  Fl e0 = fl_zero(), e2 = fl_zero(), e3 = fl_zero();
  uint64_t h = len / 2;
  for (uint64_t i = 0; i < h; i++) {
    e0 += a[i] * b[i] * c[i];
    Fl a2 = a[h + i] + a[h + i] - a[i], b2 = b[h + i] + b[h + i] - b[i], c2 = c[h + i] + c[h + i] - c[i];
    e2 += a2 * b2 * c2;
    Fl a3 = a2 + a[h + i] - a[i], b3 = b2 + b[h + i] - b[i], c3 = c2 + c[h + i] - c[i];
    e3 += a3 * b3 * c3;
  }
  memcpy(out96, e0.v, 32); memcpy(out96 + 32, e2.v, 32); memcpy(out96 + 64, e3.v, 32);
}
// SP/sparse_mlpoly.rs:467-481 (COO entries carry canonical values; z and out are Montgomery)
int orc_spmv_mont(const CooEntry *M, uint64_t nnz, uint64_t num_rows, uint64_t num_cols, const uint8_t *z, uint8_t *out) {
  FlVec zz = load_mont(z, num_cols), o(num_rows, fl_zero());
  for (uint64_t i = 0; i < nnz; i++) {
    Fl v;
    if (!fl_from_bytes(M[i].val, &v) || M[i].row >= num_rows || M[i].col >= num_cols) return 0;
    o[M[i].row] += v * zz[M[i].col];
  }
  store_mont(o, out);
  return 1;
}
// SP/sparse_mlpoly.rs:483-498
int orc_spmv_t_mont(const CooEntry *M, uint64_t nnz, uint64_t num_rows, uint64_t num_cols, const uint8_t *rx, uint8_t *out) {
  FlVec r = load_mont(rx, num_rows), o(num_cols, fl_zero());
  for (uint64_t i = 0; i < nnz; i++) {
    Fl v;
    if (!fl_from_bytes(M[i].val, &v) || M[i].row >= num_rows || M[i].col >= num_cols) return 0;
    o[M[i].col] += r[M[i].row] * v;
  }
  store_mont(o, out);
  return 1;
}
// SP/dense_mlpoly.rs:193-218 with PolyCommitmentGens::new(log2 len, label); blinds may be NULL (zeros). Z, blinds Montgomery.
void orc_hyrax_commit_mont(const uint8_t *Z, uint64_t len, const uint8_t *blinds, const char *label, int threads, uint8_t *out) {
  g_threads = threads > 0 ? threads : 1;
  DensePoly p(load_mont(Z, len));
  PolyCommitmentGens gens = pcg_new(p.num_vars, label);
  size_t l, r;
  factored_lens(p.num_vars, &l, &r);
  FlVec b = blinds ? load_mont(blinds, pow2(l)) : FlVec(pow2(l), fl_zero());
  PolyCommitment c = commit_inner(p.Z, b, gens.gens.gens_n);
  for (size_t i = 0; i < c.C.size(); i++) memcpy(out + 32 * i, c.C[i].data(), 32);
}
// SP/dense_mlpoly.rs:220-227: LZ = L * Z (Z viewed as L_size x R_size)
void orc_bound_mont(const uint8_t *Z, uint64_t len, const uint8_t *L, uint8_t *out) {
  DensePoly p(load_mont(Z, len));
  size_t l, r;
  factored_lens(p.num_vars, &l, &r);
  store_mont(dense_bound(p, load_mont(L, pow2(l))), out);
}

// ---- instances, builders and the full flow ----
void *orc_build_point_mult(uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px, const uint8_t *py) {
  return new BuiltInstance(build_point_mult(m, weights_lo_hi, px, py));
}
void *orc_build_point_add(uint64_t n, const uint8_t *px, const uint8_t *py, const uint8_t *rx, const uint8_t *ry, const int64_t *rz) {
  return new BuiltInstance(build_point_add(n, px, py, rx, ry, rz));
}
void *orc_build_custom(uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs, uint64_t nnz_param, const CooEntry *A,
                       uint64_t nA, const CooEntry *B, uint64_t nB, const CooEntry *C, uint64_t nC, const uint8_t *vars_para,
                       const uint8_t *vars_input, const uint8_t *vars, const uint8_t *inputs) {
  BuiltInstance *bi = new BuiltInstance();
  bi->num_cons = num_cons; bi->num_vars = num_vars; bi->num_inputs = num_inputs; bi->num_non_zero_entries = nnz_param;
  bi->A.assign(A, A + nA); bi->B.assign(B, B + nB); bi->C.assign(C, C + nC);
  auto ld = [](const uint8_t *p, uint64_t n) {
    std::vector<std::array<uint8_t, 32>> v(n);
    if (n) memcpy((void *)v.data(), p, 32 * n);
    return v;
  };
  bi->vars_para = ld(vars_para, num_vars); bi->vars_input = ld(vars_input, num_vars); bi->vars = ld(vars, num_vars);
  bi->inputs = ld(inputs, num_inputs);
  return bi;
}
void orc_built_info(void *h, uint64_t *out7) {
  BuiltInstance *b = (BuiltInstance *)h;
  out7[0] = b->num_cons; out7[1] = b->num_vars; out7[2] = b->num_inputs; out7[3] = b->num_non_zero_entries;
  out7[4] = b->A.size(); out7[5] = b->B.size(); out7[6] = b->C.size();
}
void orc_built_copy(void *h, CooEntry *A, CooEntry *B, CooEntry *C, uint8_t *vars_para, uint8_t *vars_input, uint8_t *vars, uint8_t *inputs) {
  BuiltInstance *b = (BuiltInstance *)h;
  memcpy((void *)A, b->A.data(), sizeof(CooEntry) * b->A.size());
  memcpy((void *)B, b->B.data(), sizeof(CooEntry) * b->B.size());
  memcpy((void *)C, b->C.data(), sizeof(CooEntry) * b->C.size());
  memcpy(vars_para, b->vars_para.data(), 32 * b->num_vars);
  memcpy(vars_input, b->vars_input.data(), 32 * b->num_vars);
  memcpy(vars, b->vars.data(), 32 * b->num_vars);
  if (b->num_inputs) memcpy(inputs, b->inputs.data(), 32 * b->num_inputs);
}
void orc_built_free(void *h) { delete (BuiltInstance *)h; }

// seeds: 32-byte canonical scalars. Returns NULL on failure (message on stderr).
void *orc_run_flow(void *h, const uint8_t *seed_q32, const uint8_t *seed_p32, int do_verify, int threads) {
  g_threads = threads > 0 ? threads : 1;
  try {
    Fl sq, sp;
    if (!fl_from_bytes(seed_q32, &sq) || !fl_from_bytes(seed_p32, &sp)) return nullptr;
    return new FlowResult(run_flow(*(BuiltInstance *)h, sq, sp, do_verify != 0));
  } catch (std::exception &e) {
    fprintf(stderr, "orc_run_flow: %s\n", e.what());
    return nullptr;
  }
}
// which: 0 proof, 1 comm, 2 comm_vars_para, 3 comm_vars_input, 4 comm_vars. Returns the length; copies if cap suffices.
uint64_t orc_flow_get(void *f, int which, uint8_t *buf, uint64_t cap) {
  FlowResult *fr = (FlowResult *)f;
  std::vector<uint8_t> tmp;
  const std::vector<uint8_t> *src = &tmp;
  if (which == 0) src = &fr->proof;
  else if (which == 1) src = &fr->comm;
  else {
    PolyCommitment &pc = which == 2 ? fr->comm_vars_para : (which == 3 ? fr->comm_vars_input : fr->comm_vars);
    for (auto &c : pc.C) tmp.insert(tmp.end(), c.begin(), c.end());
  }
  if (buf && cap >= src->size() && !src->empty()) memcpy(buf, src->data(), src->size());
  return src->size();
}
int orc_flow_verified(void *f) { return ((FlowResult *)f)->verified ? 1 : 0; }
// out: T_NTIMERS phase ms (reference timer labels) then gens, commits, verify
void orc_flow_times(void *f, double *out) {
  FlowResult *fr = (FlowResult *)f;
  for (int i = 0; i < T_NTIMERS; i++) out[i] = fr->times.ms[i];
  out[T_NTIMERS] = fr->ms_gens; out[T_NTIMERS + 1] = fr->ms_commits; out[T_NTIMERS + 2] = fr->ms_verify;
}
int orc_num_timers() { return T_NTIMERS; }
const char *orc_timer_name(int i) { return i < T_NTIMERS ? TIMER_NAMES[i] : (i == T_NTIMERS ? "gens" : (i == T_NTIMERS + 1 ? "witness_commits" : "verify")); }
void orc_flow_free(void *f) { delete (FlowResult *)f; }

// my_lib_verify on serialized artefacts (used to check proofs produced by the CUDA path).
// num_* are the UNPADDED sizes passed to SNARKGens::new; com_1 / com_2: L x 32 compressed rows.
int orc_verify(uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs, uint64_t nnz_param, const uint8_t *proof, uint64_t proof_len,
               const uint8_t *comm, uint64_t comm_len, const uint8_t *inputs, const uint8_t *com_1, const uint8_t *com_2, uint64_t L) {
  try {
    SNARK pf;
    R1CSCommitment cm;
    if (!deserialize(proof, proof_len, &pf)) return -1;
    if (!deserialize(comm, comm_len, &cm)) return -2;
    FlVec in(num_inputs);
    for (uint64_t i = 0; i < num_inputs; i++)
      if (!fl_from_bytes(inputs + 32 * i, &in[i])) return -3;
    PolyCommitment c1, c2;
    c1.C.resize(L); c2.C.resize(L);
    for (uint64_t i = 0; i < L; i++) { memcpy(c1.C[i].data(), com_1 + 32 * i, 32); memcpy(c2.C[i].data(), com_2 + 32 * i, 32); }
    SNARKGens gens = snarkgens_new(num_cons, num_vars, num_inputs, nnz_param);
    Transcript vt("snark_example");
    return my_lib_verify(pf, cm, in, vt, gens, c1, c2) ? 1 : 0;
  } catch (std::exception &e) {
    fprintf(stderr, "orc_verify: %s\n", e.what());
    return -4;
  }
}

}  // extern "C"
