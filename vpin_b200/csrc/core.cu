#include "core.cuh"

#include <algorithm>

namespace vpin {

std::atomic<uint64_t> g_kernel_launches{0};

// ------------------------------------------------------------------------------------------------ profiling
const char *prof_class_name(int cls) {
  static const char *names[PROF_COUNT] = {"msm_recode", "msm_accumulate", "msm_finish", "sumcheck_cubic_round", "sumcheck_quad_round",
                                          "sumcheck_batched_round", "bind_top", "spmv_csr", "spmv_csc", "eq_evals", "product_tree",
                                          "hash_layer", "deref_gather", "bound_LZ", "dot"};
  return cls >= 0 && cls < PROF_COUNT ? names[cls] : "?";
}
static cudaEvent_t prof_event(Ctx *c) {
  if (!c->prof.pool.empty()) { cudaEvent_t e = c->prof.pool.back(); c->prof.pool.pop_back(); return e; }
  cudaEvent_t e;
  VPIN_CUDA(cudaEventCreate(&e));
  return e;
}
ProfScope::ProfScope(Ctx *ctx, int cls_, double units, double bytes, int launches) : c(ctx), cls(cls_) {
  if (!c->prof.on || units < c->prof.min_units) return;
  Prof::Acc &a = c->prof.acc[cls];
  a.launches += launches; a.units += units; a.bytes += bytes;
  e0 = prof_event(c);
  cudaEventRecord(e0, c->st);
}
ProfScope::~ProfScope() {
  if (!e0) return;
  cudaEvent_t e1 = prof_event(c);
  cudaEventRecord(e1, c->st);
  c->prof.pending.push_back({cls, e0, e1});
}
void prof_drain(Ctx *c) {
  c->sync();
  for (auto &p : c->prof.pending) {
    float ms = 0;
    cudaEventElapsedTime(&ms, p.e0, p.e1);
    c->prof.acc[p.cls].ms += ms;
    c->prof.pool.push_back(p.e0);
    c->prof.pool.push_back(p.e1);
  }
  c->prof.pending.clear();
}

// ------------------------------------------------------------------------------------------------ host fixed base
void HostBase::build(const ge_t &p) {
  const int P = 64, M = 8;
  std::vector<ge_t> ext(P * M);
  ge_t base = p;
  for (int pos = 0; pos < P; pos++) {
    ge_t cur = base;
    for (int m = 0; m < M; m++) {
      ext[pos * M + m] = cur;
      if (m + 1 < M) cur = ge_add(cur, base);
    }
    for (int k = 0; k < 4; k++) base = ge_dbl(base);
  }
  // batch inversion of all Z
  std::vector<fp_t> prefix(P * M);
  fp_t run = fp_one();
  for (int i = 0; i < P * M; i++) { run = fp_mul(run, ext[i].Z); prefix[i] = run; }
  fp_t inv = fp_invert(run);
  tbl.resize(P * M);
  for (int i = P * M - 1; i >= 0; i--) {
    fp_t zinv = i > 0 ? fp_mul(inv, prefix[i - 1]) : inv;
    inv = fp_mul(inv, ext[i].Z);
    tbl[i] = ge_to_niels(ext[i], zinv);
  }
}
void HostBase::mul_acc(const fl_t &s_mont, ge_t *acc) const {
  fl_t s = fl_from_mont(s_mont);
  int carry = 0;
  for (int pos = 0; pos < 64; pos++) {
    int d = (int)((s.v[pos >> 3] >> ((pos & 7) * 4)) & 15u) + carry;
    carry = 0;
    if (d > 8) { d -= 16; carry = 1; }
    if (d > 0) *acc = ge_madd(*acc, tbl[pos * 8 + d - 1]);
    else if (d < 0) *acc = ge_msub(*acc, tbl[pos * 8 - d - 1]);
  }
}
ge_t HostBase::mul(const fl_t &s_mont) const {
  ge_t acc = ge_identity();
  mul_acc(s_mont, &acc);
  return acc;
}

// ------------------------------------------------------------------------------------------------ generators
std::shared_ptr<LabelGens> get_label_gens(Ctx *ctx, const std::string &label, size_t n) {
  auto it = ctx->label_gens.find(label);
  if (it != ctx->label_gens.end() && it->second->n >= n) return it->second;
  auto g = std::make_shared<LabelGens>();
  g->label = label;
  g->n = n;
  // SHAKE256(label || compressed basepoint) read as one stream of 64-byte blocks (Spartan/src/commitments.rs:21-31)
  static const uint8_t kBasepoint[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9,
                                         0x61, 0xc5, 0x00, 0x51, 0x5f, 0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82,
                                         0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
  Shake256Xof xof;
  xof.update(label.data(), label.size());
  xof.update(kBasepoint, 32);
  std::vector<uint8_t> stream(n * 64);
  xof.read(stream.data(), stream.size());
  DevVec<uint8_t> d_stream(stream.size(), ctx->st);
  d_stream.upload(stream.data(), stream.size());
  g->d_pts.alloc(n, ctx->st);
  launch_from_uniform_bytes(d_stream.p, n, g->d_pts.p, ctx->st);
  g->h_pts.resize(n);
  g->d_pts.download(g->h_pts.data(), n);
  g->d_table.alloc(msm_table_entries(n), ctx->st);
  {
    DevVec<ge_t> scratch(n, ctx->st);
    launch_table_build(g->d_pts.p, n, g->d_table.p, scratch.p, ctx->st);
  }
  ctx->sync();
  ctx->label_gens[label] = g;
  return g;
}

// ------------------------------------------------------------------------------------------------ Hyrax rows
void hyrax_rows(Ctx *ctx, const LabelGens &g, const fl_t *dZ, size_t rows, size_t cols, size_t ld, const fl_t *d_blinds,
                size_t blind_base, ge_t *d_points, uint8_t *d_comp) {
  VPIN_REQUIRE(cols <= g.n && (!d_blinds || blind_base < g.n), VPIN_ERR_SIZE_MISMATCH, "hyrax_rows: not enough generators");
  size_t cols_total = cols + (d_blinds ? 1 : 0);
  size_t stride = msm_col_stride(cols_total);
  // bound the digit buffer (2 bytes x windows per scalar) to ~1 GiB per pass
  size_t max_rows = ((size_t)1 << 30) / (stride * kMsmWindows * sizeof(uint16_t));
  if (max_rows < 1) max_rows = 1;
  size_t chunk = rows < max_rows ? rows : max_rows;
  size_t segs = msm_num_segments(chunk, cols_total);
  DevVec<uint16_t> digits(msm_digits_count(chunk, cols_total), ctx->st);
  DevVec<ge_t> partial(chunk * kMsmGroup * segs, ctx->st), sums(segs > 1 ? chunk * kMsmGroup : 0, ctx->st);
  for (size_t r0 = 0; r0 < rows; r0 += chunk) {
    size_t nr = rows - r0 < chunk ? rows - r0 : chunk;
    double pts = (double)nr * cols_total;
    {
      ProfScope ps(ctx, PROF_MSM_RECODE, pts, pts * (32 + 2 * kMsmWindows));
      launch_recode(dZ + r0 * ld, nr, cols, ld, d_blinds ? d_blinds + r0 : nullptr, digits.p, ctx->d_counters.p, ctx->st);
    }
    {
      ProfScope ps(ctx, PROF_MSM_ACCUMULATE, pts, 0);
      launch_msm_accumulate(g.table(), digits.p, nr, cols, d_blinds != nullptr, blind_base, segs, partial.p, ctx->st);
    }
    ProfScope ps(ctx, PROF_MSM_FINISH, pts, 0);
    launch_msm_finish(partial.p, nr, segs, sums.p, d_points ? d_points + r0 : nullptr, d_comp ? d_comp + 32 * r0 : nullptr, ctx->st);
  }
}

// ------------------------------------------------------------------------------------------------ instance
static void build_matrix(Ctx *ctx, MatrixDev &m, const std::vector<uint32_t> &row, const std::vector<uint32_t> &col,
                         const std::vector<fl_t> &val, size_t num_rows, size_t num_cols) {
  size_t nnz = row.size();
  m.nnz = nnz;
  m.h_row = row;
  m.h_col = col;
  cudaStream_t st = ctx->st;
  m.coo_row.alloc(nnz, st); m.coo_col.alloc(nnz, st); m.coo_val.alloc(nnz, st);
  if (nnz) { m.coo_row.upload(row.data(), nnz); m.coo_col.upload(col.data(), nnz); m.coo_val.upload(val.data(), nnz); }
  // CSR by counting sort on rows
  std::vector<uint32_t> ptr(num_rows + 1, 0), idx(nnz);
  std::vector<fl_t> v(nnz);
  for (size_t i = 0; i < nnz; i++) ptr[row[i] + 1]++;
  for (size_t r = 0; r < num_rows; r++) ptr[r + 1] += ptr[r];
  {
    std::vector<uint32_t> pos(ptr.begin(), ptr.end() - 1);
    for (size_t i = 0; i < nnz; i++) { uint32_t p = pos[row[i]]++; idx[p] = col[i]; v[p] = val[i]; }
  }
  m.csr_ptr.alloc(num_rows + 1, st); m.csr_col.alloc(nnz, st); m.csr_val.alloc(nnz, st);
  m.csr_ptr.upload(ptr.data(), num_rows + 1);
  if (nnz) { m.csr_col.upload(idx.data(), nnz); m.csr_val.upload(v.data(), nnz); }
  ctx->sync();
  // CSC by counting sort on columns
  std::vector<uint32_t> cptr(num_cols + 1, 0);
  for (size_t i = 0; i < nnz; i++) cptr[col[i] + 1]++;
  for (size_t c = 0; c < num_cols; c++) cptr[c + 1] += cptr[c];
  {
    std::vector<uint32_t> pos(cptr.begin(), cptr.end() - 1);
    for (size_t i = 0; i < nnz; i++) { uint32_t p = pos[col[i]]++; idx[p] = row[i]; v[p] = val[i]; }
  }
  std::vector<uint32_t> longc;
  for (size_t c = 0; c < num_cols; c++)
    if (cptr[c + 1] - cptr[c] > (uint32_t)kLongCol) longc.push_back((uint32_t)c);
  m.n_long = longc.size();
  m.csc_ptr.alloc(num_cols + 1, st); m.csc_row.alloc(nnz, st); m.csc_val.alloc(nnz, st); m.long_cols.alloc(longc.size(), st);
  m.csc_ptr.upload(cptr.data(), num_cols + 1);
  if (nnz) { m.csc_row.upload(idx.data(), nnz); m.csc_val.upload(v.data(), nnz); }
  if (!longc.empty()) m.long_cols.upload(longc.data(), longc.size());
  ctx->sync();
}

// Spartan/src/lib.rs:138-244 (Instance::new): padding rules and error behaviour; the zlib digest (:241) is unused on
// the SNARK path and is not computed.
std::unique_ptr<Instance> instance_create(Ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs,
                                          const vpin_coo_entry *A, uint64_t nA, const vpin_coo_entry *B, uint64_t nB,
                                          const vpin_coo_entry *C, uint64_t nC) {
  size_t num_vars_padded = next_pow2(std::max<size_t>(num_vars, num_inputs + 1));
  size_t num_cons_padded = num_cons;
  if (num_cons_padded == 0 || num_cons_padded == 1) num_cons_padded = 2;
  if (next_pow2(num_cons) != num_cons) num_cons_padded = next_pow2(num_cons);
  VPIN_REQUIRE(2 * num_vars_padded < ((size_t)1 << 32) && num_cons_padded < ((size_t)1 << 32), VPIN_ERR_BAD_ARGUMENT,
               "instance too large for 32-bit indices");
  auto inst = std::make_unique<Instance>();
  inst->num_cons = num_cons_padded;
  inst->num_vars = num_vars_padded;
  inst->num_inputs = num_inputs;
  const vpin_coo_entry *src[3] = {A, B, C};
  uint64_t cnt[3] = {nA, nB, nC};
  for (int k = 0; k < 3; k++) {
    std::vector<uint32_t> row, col;
    std::vector<fl_t> val;
    row.reserve(cnt[k]); col.reserve(cnt[k]); val.reserve(cnt[k]);
    for (uint64_t i = 0; i < cnt[k]; i++) {
      const vpin_coo_entry &e = src[k][i];
      VPIN_REQUIRE(e.row < num_cons, VPIN_ERR_INVALID_INDEX, "InvalidIndex: row");
      VPIN_REQUIRE(e.col < num_vars + 1 + num_inputs, VPIN_ERR_INVALID_INDEX, "InvalidIndex: col");
      fl_t v;
      VPIN_REQUIRE(fl_from_bytes(e.val, &v), VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
      row.push_back((uint32_t)e.row);
      col.push_back((uint32_t)(e.col >= num_vars ? e.col + num_vars_padded - num_vars : e.col));
      val.push_back(v);
    }
    if (num_cons == 0 || num_cons == 1)
      for (size_t i = cnt[k]; i < num_cons_padded; i++) { row.push_back((uint32_t)i); col.push_back((uint32_t)num_vars); val.push_back(fl_zero()); }
    build_matrix(ctx, inst->M[k], row, col, val, num_cons_padded, 2 * num_vars_padded);
  }
  return inst;
}

}  // namespace vpin
