// extern "C" surface of libvpin_b200.so (include/vpin_b200.h). Thin: argument checks, byte-format conversion,
// exception -> status mapping. The prover itself is in prover.cu, the kernels in kernels_*.cu.
#include "prover.cuh"

#include <errno.h>
#include <sys/random.h>

using namespace vpin;

#define VPIN_TRY(ctx_)                      \
  Ctx *c_ = (ctx_);                         \
  if (c_) cudaSetDevice(c_->device); /* contexts may be driven from different host threads */ \
  LaunchCounter::current() = c_ ? &c_->kernel_launches : nullptr; /* launches of this call count for this context */ \
  try {
#define VPIN_CATCH                                                    \
  }                                                                   \
  catch (const vpin::Error &e) {                                      \
    if (c_) c_->err = e.what();                                       \
    return e.code;                                                    \
  }                                                                   \
  catch (const std::bad_alloc &) {                                    \
    if (c_) c_->err = "host allocation failed";                       \
    return VPIN_ERR_OOM;                                              \
  }                                                                   \
  catch (const std::exception &e) {                                   \
    if (c_) c_->err = e.what();                                       \
    return VPIN_ERR_PROVER;                                           \
  }                                                                   \
  return VPIN_OK;

namespace {
// true when a pre-launched round kernel of the context gave up waiting for its challenge (kernels_poly.cuh ChalLatch::abort)
bool mailbox_timed_out(Ctx *ctx) {
  uint32_t flag = 0;
  if (!ctx->d_chal_latch.p || cudaStreamSynchronize(ctx->st) != cudaSuccess) return false;
  if (cudaMemcpy(&flag, &ctx->d_chal_latch.p->abort, sizeof(flag), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
  return flag != 0;
}
// canonical LE bytes -> Montgomery table in HBM (rejects values >= l like Scalar::from_bytes)
DevVec<fl_t> upload_scalars(Ctx *ctx, const uint8_t *b, size_t n) {
  DevVec<fl_t> d(n, ctx->st);
  if (!n) return d;
  // raw canonical bytes go up as they are; the Montgomery conversion and the `>= l` check run on the device
  VPIN_CUDA(cudaMemcpyAsync(d.p, b, n * sizeof(fl_t), cudaMemcpyHostToDevice, ctx->st));
  uint32_t *d_bad = reinterpret_cast<uint32_t *>(ctx->d_counters.p + 1);
  VPIN_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(uint32_t), ctx->st));
  launch_from_bytes_checked(d.p, n, d.p, d_bad, ctx->st);
  uint32_t bad = 0;
  VPIN_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->st));
  ctx->sync();
  VPIN_REQUIRE(bad == 0, VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
  return d;
}
void download_scalars(Ctx *ctx, const fl_t *d, size_t n, uint8_t *out) {
  if (!n) return;
  DevVec<fl_t> tmp(n, ctx->st);
  launch_from_mont(d, n, tmp.p, ctx->st);
  VPIN_CUDA(cudaMemcpyAsync(out, tmp.p, n * sizeof(fl_t), cudaMemcpyDeviceToHost, ctx->st));
  ctx->sync();
}
bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
}  // namespace

extern "C" {

vpin_status vpin_ctx_create(int32_t cuda_device, vpin_ctx **out) { return vpin_ctx_create_ex(cuda_device, 0, out); }
vpin_status vpin_ctx_create_ex(int32_t cuda_device, int32_t high_priority, vpin_ctx **out) {
  if (!out) return VPIN_ERR_BAD_ARGUMENT;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cuda_device < 0 || cuda_device >= ndev) return VPIN_ERR_CUDA;
  Ctx *ctx = new Ctx();
  try {
    ctx->device = cuda_device;
    VPIN_CUDA(cudaSetDevice(cuda_device));
    {
      int lo = 0, hi = 0;  // (numerically lower = more urgent; default streams sit at `lo`)
      VPIN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      // (a default context sits one step above the lowest priority: its side stream - SideScope - runs below it)
      ctx->background = high_priority < 0;
      VPIN_CUDA(cudaStreamCreateWithPriority(&ctx->st, cudaStreamNonBlocking,
                                             high_priority > 0 ? hi : (high_priority < 0 ? lo : (lo - 1 >= hi ? lo - 1 : lo))));
    }
    block_cache_register(ctx->st);
    block_cache_set_pressure_hook(ctx->st, [ctx]() {
      if (ctx->workspace_busy || !ctx->workspace.p) return false;
      ctx->workspace.release();  // back to the cache's idle list; the allocator returns that to the driver next
      return true;
    });
    ctx->d_partials.alloc((size_t)3 * kRedBlocks * 32, ctx->st);
    ctx->d_small.alloc(256, ctx->st);
    ctx->d_counters.alloc(4, ctx->st);
    ctx->d_counters.zero();
    VPIN_CUDA(cudaMallocHost((void **)&ctx->h_small, 512 * sizeof(fl_t)));
    VPIN_CUDA(cudaHostAlloc((void **)&ctx->h_slots, 2 * sizeof(RoundSlot), cudaHostAllocMapped));
    memset(ctx->h_slots, 0, 2 * sizeof(RoundSlot));
    VPIN_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_slots, ctx->h_slots, 0));
    VPIN_CUDA(cudaHostAlloc((void **)&ctx->h_tail, kTailElems * sizeof(fl_t), cudaHostAllocMapped));
    VPIN_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_tail, ctx->h_tail, 0));
    ctx->d_round_counters.alloc(1 + kMaxBatched + 13, ctx->st);
    ctx->d_round_counters.zero();
    VPIN_CUDA(cudaHostAlloc((void **)&ctx->h_chal, kChalRing * sizeof(ChalSlot), cudaHostAllocMapped));
    memset(ctx->h_chal, 0, kChalRing * sizeof(ChalSlot));
    VPIN_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_chal, ctx->h_chal, 0));
    ctx->d_chal_latch.alloc(1, ctx->st);
    ctx->d_chal_latch.zero();
    ctx->sync();
  } catch (const std::exception &) {
    delete ctx;
    return VPIN_ERR_CUDA;
  }
  *out = reinterpret_cast<vpin_ctx *>(ctx);
  return VPIN_OK;
}
void vpin_ctx_destroy(vpin_ctx *ctx_) {
  Ctx *ctx = reinterpret_cast<Ctx *>(ctx_);
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  try { dist_destroy(ctx); } catch (...) {}
  ctx->label_gens.clear();
  ctx->d_partials.release();
  ctx->d_small.release();
  ctx->d_counters.release();
  for (auto &p : ctx->prof.pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
  for (auto e : ctx->prof.pool) cudaEventDestroy(e);
  if (ctx->ev_marker) cudaEventDestroy(ctx->ev_marker);
  if (ctx->h_small) gated_cuda_free_host(ctx->h_small);
  if (ctx->h_slots) gated_cuda_free_host(ctx->h_slots);
  if (ctx->h_tail) gated_cuda_free_host(ctx->h_tail);
  if (ctx->h_chal) gated_cuda_free_host(ctx->h_chal);
  ctx->d_chal_latch.release();
  ctx->d_round_counters.release();
  ctx->workspace.release();
  cudaStreamSynchronize(ctx->st);
  if (ctx->st_side) {
    cudaStreamSynchronize(ctx->st_side);
    block_cache_unregister(ctx->st_side);
    cudaStreamDestroy(ctx->st_side);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
  }
  block_cache_unregister(ctx->st);
  cudaStreamDestroy(ctx->st);
  delete ctx;
}
vpin_status vpin_nccl_unique_id(uint8_t id_out[128]) {
  if (!id_out) return VPIN_ERR_BAD_ARGUMENT;
  try {
    dist_get_unique_id(id_out);
  } catch (const std::exception &) {
    return VPIN_ERR_CUDA;
  }
  return VPIN_OK;
}
vpin_status vpin_ctx_init_distributed(vpin_ctx *ctx, int32_t rank, int32_t world, const uint8_t nccl_id[128]) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(world == 1 || nccl_id, VPIN_ERR_BAD_ARGUMENT, "null id");
  dist_init(c_, rank, world, nccl_id);
  VPIN_CATCH
}
vpin_status vpin_ctx_set_shard_sumcheck(vpin_ctx *ctx, int32_t on) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(c_, VPIN_ERR_BAD_ARGUMENT, "null context");
  c_->shard_sumcheck = on < 0 ? -1 : (on ? 1 : 0);
  VPIN_CATCH
}
void vpin_shard_rows(uint64_t rows, int32_t rank, int32_t world, uint64_t *r0, uint64_t *r1, int32_t *sharded) {
  size_t a, b;
  bool s = shard_rows(rows, rank, world, &a, &b);
  if (r0) *r0 = a;
  if (r1) *r1 = b;
  if (sharded) *sharded = s ? 1 : 0;
}
const char *vpin_last_error(const vpin_ctx *ctx) { return ctx ? reinterpret_cast<const Ctx *>(ctx)->err.c_str() : "null context"; }
uint64_t vpin_kernel_launches(const vpin_ctx *ctx) {  // launches made by calls on THIS context (process total when ctx is NULL)
  return ctx ? reinterpret_cast<const Ctx *>(ctx)->kernel_launches.load() : g_kernel_launches.load();
}
void *vpin_stream(vpin_ctx *ctx) { return reinterpret_cast<Ctx *>(ctx)->st; }
vpin_status vpin_sync(vpin_ctx *ctx) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  c_->sync();
  VPIN_CATCH
}

// ---------------------------------------------------------------------------------------------- gens / instance
vpin_status vpin_gens_create(vpin_ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs, uint64_t num_nz_entries,
                             vpin_gens **out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(out, VPIN_ERR_BAD_ARGUMENT, "null out");
  *out = reinterpret_cast<vpin_gens *>(snark_gens_create(c_, num_cons, num_vars, num_inputs, num_nz_entries).release());
  VPIN_CATCH
}
void vpin_gens_destroy(vpin_gens *g) { delete reinterpret_cast<SnarkGens *>(g); }
vpin_status vpin_gens_witness_grid(const vpin_gens *g, uint64_t *L, uint64_t *R) {
  if (!g || !L || !R) return VPIN_ERR_BAD_ARGUMENT;
  const SnarkGens *sg = reinterpret_cast<const SnarkGens *>(g);
  *L = sg->sat_pc.L;
  *R = sg->sat_pc.R;
  return VPIN_OK;
}
vpin_status vpin_instance_create(vpin_ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs, const vpin_coo_entry *A,
                                 uint64_t nA, const vpin_coo_entry *B, uint64_t nB, const vpin_coo_entry *C, uint64_t nC,
                                 vpin_instance **out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(out, VPIN_ERR_BAD_ARGUMENT, "null out");
  *out = reinterpret_cast<vpin_instance *>(instance_create(c_, num_cons, num_vars, num_inputs, A, nA, B, nB, C, nC).release());
  VPIN_CATCH
}
void vpin_instance_destroy(vpin_instance *inst) { delete reinterpret_cast<Instance *>(inst); }
vpin_status vpin_instance_dims(const vpin_instance *inst, uint64_t *nc, uint64_t *nv, uint64_t *ni) {
  if (!inst) return VPIN_ERR_BAD_ARGUMENT;
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  if (nc) *nc = I->num_cons;
  if (nv) *nv = I->num_vars;
  if (ni) *ni = I->num_inputs;
  return VPIN_OK;
}
vpin_status vpin_instance_is_sat(vpin_ctx *ctx, const vpin_instance *inst, const uint8_t *vars32, uint64_t n_vars,
                                 const uint8_t *inputs32, uint64_t n_inputs, int32_t *sat) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  VPIN_REQUIRE(I && sat, VPIN_ERR_BAD_ARGUMENT, "null argument");
  VPIN_REQUIRE(n_vars <= I->num_vars && n_inputs == I->num_inputs, VPIN_ERR_INVALID_NUM_INPUTS, "InvalidNumberOfInputs");
  *sat = instance_is_sat(c_, *I, vars32, n_vars, inputs32, n_inputs) ? 1 : 0;
  VPIN_CATCH
}

// COO triples of the instance in Instance::new's format (unpadded column indices, canonical values): lets a caller that
// built the instance with vpin_build_point_* replay Instance::new from host buffers (bench.py's e2e leg).
vpin_status vpin_instance_nnz(const vpin_instance *inst, uint64_t nnz_out[3]) {
  if (!inst || !nnz_out) return VPIN_ERR_BAD_ARGUMENT;
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  for (int k = 0; k < 3; k++) nnz_out[k] = I->M[k].nnz;
  return VPIN_OK;
}
vpin_status vpin_instance_export_coo(vpin_ctx *ctx, const vpin_instance *inst, uint64_t num_vars_unpadded, int32_t which,
                                     vpin_coo_entry *out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  VPIN_REQUIRE(I && out && which >= 0 && which < 3 && num_vars_unpadded <= I->num_vars, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  const MatrixDev &m = I->M[which];
  std::vector<fl_t> v(m.nnz);
  std::vector<uint32_t> h_row(m.nnz), h_col(m.nnz);
  if (m.nnz) {
    DevVec<fl_t> tmp(m.nnz, c_->st);
    launch_from_mont(m.coo_val.p, m.nnz, tmp.p, c_->st);
    tmp.download(v.data(), m.nnz);
    m.coo_row.download(h_row.data(), m.nnz);
    m.coo_col.download(h_col.data(), m.nnz);
    c_->sync();
  }
  for (size_t i = 0; i < m.nnz; i++) {
    out[i].row = h_row[i];
    uint64_t col = h_col[i];
    out[i].col = col >= I->num_vars ? col - (I->num_vars - num_vars_unpadded) : col;  // undo the shift of Spartan/src/lib.rs:196-200
    memcpy(out[i].val, v[i].v, 32);
  }
  VPIN_CATCH
}

vpin_status vpin_encode(vpin_ctx *ctx, const vpin_instance *inst, const vpin_gens *gens, uint8_t *comm_out, uint64_t comm_cap,
                        uint64_t *comm_len, vpin_decomm **decomm) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(inst && gens && comm_len && decomm, VPIN_ERR_BAD_ARGUMENT, "null argument");
  std::vector<uint8_t> comm;
  auto d = snark_encode(c_, *reinterpret_cast<const Instance *>(inst), *reinterpret_cast<const SnarkGens *>(gens), &comm);
  *comm_len = comm.size();
  VPIN_REQUIRE(comm_out && comm_cap >= comm.size(), VPIN_ERR_BUFFER_TOO_SMALL, "comm_out too small");
  memcpy(comm_out, comm.data(), comm.size());
  *decomm = reinterpret_cast<vpin_decomm *>(d.release());
  VPIN_CATCH
}
vpin_status vpin_encode_tables(vpin_ctx *ctx, const vpin_instance *inst, const vpin_gens *gens, vpin_decomm **decomm) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(inst && gens && decomm, VPIN_ERR_BAD_ARGUMENT, "null argument");
  auto d = snark_encode_tables(c_, *reinterpret_cast<const Instance *>(inst), *reinterpret_cast<const SnarkGens *>(gens));
  *decomm = reinterpret_cast<vpin_decomm *>(d.release());
  VPIN_CATCH
}
vpin_status vpin_encode_commit(vpin_ctx *ctx, const vpin_decomm *decomm, const vpin_gens *gens, uint8_t *comm_out, uint64_t comm_cap,
                               uint64_t *comm_len) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(decomm && gens && comm_len, VPIN_ERR_BAD_ARGUMENT, "null argument");
  std::vector<uint8_t> comm = snark_encode_commit(c_, *reinterpret_cast<const Decomm *>(decomm), *reinterpret_cast<const SnarkGens *>(gens));
  *comm_len = comm.size();
  VPIN_REQUIRE(comm_out && comm_cap >= comm.size(), VPIN_ERR_BUFFER_TOO_SMALL, "comm_out too small");
  memcpy(comm_out, comm.data(), comm.size());
  VPIN_CATCH
}
void vpin_decomm_destroy(vpin_decomm *d) { delete reinterpret_cast<Decomm *>(d); }

// ---------------------------------------------------------------------------------------------- commitments
// init_randomness of a RandomTape (SP/random.rs:16-18): the caller's 32 canonical bytes, or - for NULL - 64 bytes from the
// operating system's CSPRNG reduced mod l, which is what the reference does (Scalar::random(&mut OsRng)). A tape seed must be
// used for ONE proof: the blinds and sigma-protocol nonces derive from it, and repeating them under another witness leaks it.
static bool tape_seed_from(const uint8_t *seed32, fl_t *out) {
  if (seed32) return fl_from_bytes(seed32, out);
  uint8_t wide[64];
  size_t got = 0;
  while (got < sizeof(wide)) {
    ssize_t r = getrandom(wide + got, sizeof(wide) - got, 0);
    if (r < 0) {
      if (errno == EINTR) continue;
      throw Error(VPIN_ERR_PROVER, "getrandom failed: no randomness for the prover's tape");
    }
    got += (size_t)r;
  }
  *out = fl_from_bytes_wide(wide);
  return true;
}
vpin_status vpin_tape_init(uint8_t tape_state[256], const uint8_t *name, uint64_t name_len, const uint8_t init_randomness32[32]) {
  static_assert(sizeof(ProverTape) <= 256, "tape state too small");
  if (!tape_state) return VPIN_ERR_BAD_ARGUMENT;
  fl_t seed;
  try {
    if (!tape_seed_from(init_randomness32, &seed)) return VPIN_ERR_INVALID_SCALAR;
  } catch (const vpin::Error &e) {
    return e.code;
  }
  ProverTape t(name, name_len, seed);
  memset(tape_state, 0, 256);
  memcpy(tape_state, &t, sizeof(t));
  return VPIN_OK;
}
vpin_status vpin_poly_commit(vpin_ctx *ctx, const vpin_gens *gens, const uint8_t *Z32, uint64_t n, uint8_t *tape_state,
                             uint8_t *points_out, uint8_t *blinds_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const SnarkGens *sg = reinterpret_cast<const SnarkGens *>(gens);
  VPIN_REQUIRE(sg && Z32 && points_out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  VPIN_REQUIRE(n == sg->sat_pc.L * sg->sat_pc.R, VPIN_ERR_SIZE_MISMATCH, "polynomial size does not match gens");
  size_t L = sg->sat_pc.L;
  std::vector<fl_t> blinds(L, fl_zero());
  if (tape_state) {
    ProverTape t(nullptr, 0, fl_zero());
    memcpy(&t, tape_state, sizeof(t));
    blinds = t.vector("poly_blinds", L);  // Spartan/src/dense_mlpoly.rs:207-210
    memcpy(tape_state, &t, sizeof(t));
  }
  DevVec<fl_t> dZ = upload_scalars(c_, Z32, n);
  DevVec<fl_t> dB(L, c_->st);
  dB.upload(blinds.data(), L);
  DevVec<uint8_t> dC(32 * L, c_->st);
  hyrax_rows(c_, *sg->sat_label, dZ.p, L, sg->sat_pc.R, sg->sat_pc.R, tape_state ? dB.p : nullptr, sg->sat_pc.h_index, nullptr, dC.p);
  dC.download(points_out, 32 * L);
  c_->sync();
  if (blinds_out)
    for (size_t i = 0; i < L; i++) fl_to_bytes(blinds[i], blinds_out + 32 * i);
  VPIN_CATCH
}
vpin_status vpin_poly_commit_with_blinds(vpin_ctx *ctx, const vpin_gens *gens, const uint8_t *Z32, uint64_t n, const uint8_t *blind1_32,
                                         const uint8_t *blind2_32, uint64_t L, uint8_t *points_out, uint8_t *blinds_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const SnarkGens *sg = reinterpret_cast<const SnarkGens *>(gens);
  VPIN_REQUIRE(sg && Z32 && points_out && blind1_32 && blind2_32, VPIN_ERR_BAD_ARGUMENT, "null argument");
  VPIN_REQUIRE(n == sg->sat_pc.L * sg->sat_pc.R && L == sg->sat_pc.L, VPIN_ERR_SIZE_MISMATCH, "polynomial size does not match gens");
  std::vector<fl_t> blinds(L);
  for (size_t i = 0; i < L; i++) {
    fl_t a, b;
    VPIN_REQUIRE(fl_from_bytes(blind1_32 + 32 * i, &a) && fl_from_bytes(blind2_32 + 32 * i, &b), VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
    blinds[i] = fl_add(a, b);  // vPIN_proof_generation/src/commit_test.rs:44-47
  }
  DevVec<fl_t> dZ = upload_scalars(c_, Z32, n);
  DevVec<fl_t> dB(L, c_->st);
  dB.upload(blinds.data(), L);
  DevVec<uint8_t> dC(32 * L, c_->st);
  hyrax_rows(c_, *sg->sat_label, dZ.p, L, sg->sat_pc.R, sg->sat_pc.R, dB.p, sg->sat_pc.h_index, nullptr, dC.p);
  dC.download(points_out, 32 * L);
  c_->sync();
  if (blinds_out)
    for (size_t i = 0; i < L; i++) fl_to_bytes(blinds[i], blinds_out + 32 * i);
  VPIN_CATCH
}
vpin_status vpin_commitments_add(vpin_ctx *ctx, const uint8_t *c1, const uint8_t *c2, uint64_t L, uint8_t *out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(c1 && c2 && out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  DevVec<uint8_t> d1(32 * L, c_->st), d2(32 * L, c_->st), ok(2 * L, c_->st), dout(32 * L, c_->st);
  DevVec<ge_t> p1(L, c_->st), p2(L, c_->st);
  d1.upload(c1, 32 * L);
  d2.upload(c2, 32 * L);
  launch_decompress(d1.p, L, p1.p, ok.p, c_->st);
  launch_decompress(d2.p, L, p2.p, ok.p + L, c_->st);
  launch_points_add(p1.p, p2.p, L, p1.p, c_->st);
  launch_compress(p1.p, L, dout.p, c_->st);
  std::vector<uint8_t> hok(2 * L);
  ok.download(hok.data(), 2 * L);
  dout.download(out, 32 * L);
  c_->sync();
  for (uint8_t v : hok) VPIN_REQUIRE(v == 1, VPIN_ERR_BAD_ARGUMENT, "point does not decompress");
  VPIN_CATCH
}

// ---------------------------------------------------------------------------------------------- prove
vpin_status vpin_witness_upload(vpin_ctx *ctx, const vpin_gens *gens, const uint8_t *vars32, uint64_t n_vars,
                                const uint8_t *comm_vars_points, const uint8_t *blinds_vars32, uint64_t L, vpin_witness **out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const SnarkGens *sg = reinterpret_cast<const SnarkGens *>(gens);
  VPIN_REQUIRE(sg && vars32 && comm_vars_points && blinds_vars32 && out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  VPIN_REQUIRE(n_vars == sg->sat_pc.L * sg->sat_pc.R && L == sg->sat_pc.L, VPIN_ERR_SIZE_MISMATCH, "witness size does not match gens");
  auto w = std::make_unique<Witness>();
  w->n_vars = n_vars;
  w->d_vars = upload_scalars(c_, vars32, n_vars);
  w->d_blinds = upload_scalars(c_, blinds_vars32, L);
  w->comm.assign(comm_vars_points, comm_vars_points + 32 * L);
  *out = reinterpret_cast<vpin_witness *>(w.release());
  VPIN_CATCH
}
void vpin_witness_destroy(vpin_witness *w) { delete reinterpret_cast<Witness *>(w); }
vpin_status vpin_prove_resident(vpin_ctx *ctx, const vpin_instance *inst, const vpin_decomm *decomm, const vpin_witness *w,
                                const uint8_t *inputs32, uint64_t n_inputs, const vpin_gens *gens, const uint8_t *transcript_label,
                                uint64_t label_len, const uint8_t tape_seed32[32], uint8_t *proof_out, uint64_t proof_cap,
                                uint64_t *proof_len) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(inst && decomm && w && gens && transcript_label && proof_len, VPIN_ERR_BAD_ARGUMENT, "null argument");
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  VPIN_REQUIRE(n_inputs == I->num_inputs, VPIN_ERR_INVALID_NUM_INPUTS, "InvalidNumberOfInputs");
  std::vector<fl_t> inputs(n_inputs);
  for (size_t i = 0; i < n_inputs; i++) VPIN_REQUIRE(fl_from_bytes(inputs32 + 32 * i, &inputs[i]), VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
  fl_t seed;
  VPIN_REQUIRE(tape_seed_from(tape_seed32, &seed), VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
  // A pre-launched round kernel whose challenge never came (its host thread was held up for seconds: a device-synchronising call
  // of another library in the process, a stopped process) times out, and the proof fails with a drained stream. The context then
  // stops pre-launching and the proof is redone once - same inputs, same tape seed, hence the same bytes.
  std::vector<uint8_t> proof;
  for (int attempt = 0;; attempt++) {
    try {
      proof = snark_prove(c_, *I, *reinterpret_cast<const Decomm *>(decomm), *reinterpret_cast<const Witness *>(w), inputs,
                          *reinterpret_cast<const SnarkGens *>(gens), transcript_label, label_len, seed);
      break;
    } catch (const vpin::Error &) {
      // (not on a distributed context: the other ranks are half-way through the collectives of the first attempt - there the
      // time-out is ten times longer and a time-out fails the call on this rank like a failed peer does on the others)
      if (attempt > 0 || c_->no_prelaunch || c_->world > 1 || !mailbox_timed_out(c_)) throw;
      c_->no_prelaunch = true;
    }
  }
  *proof_len = proof.size();
  VPIN_REQUIRE(proof_out && proof_cap >= proof.size(), VPIN_ERR_BUFFER_TOO_SMALL, "proof_out too small");
  memcpy(proof_out, proof.data(), proof.size());
  VPIN_CATCH
}
vpin_status vpin_prove(vpin_ctx *ctx, const vpin_instance *inst, const vpin_decomm *decomm, const uint8_t *vars32, uint64_t n_vars,
                       const uint8_t *inputs32, uint64_t n_inputs, const vpin_gens *gens, const uint8_t *transcript_label,
                       uint64_t label_len, const uint8_t *comm_vars_points, const uint8_t *blinds_vars32, uint64_t L,
                       const uint8_t tape_seed32[32], uint8_t *proof_out, uint64_t proof_cap, uint64_t *proof_len) {
  vpin_witness *w = nullptr;
  vpin_status s = vpin_witness_upload(ctx, gens, vars32, n_vars, comm_vars_points, blinds_vars32, L, &w);
  if (s != VPIN_OK) return s;
  s = vpin_prove_resident(ctx, inst, decomm, w, inputs32, n_inputs, gens, transcript_label, label_len, tape_seed32, proof_out, proof_cap,
                          proof_len);
  vpin_witness_destroy(w);
  return s;
}
uint32_t vpin_last_phase_times(const vpin_ctx *ctx, const char **names_out, double *ms_out, uint32_t cap) {
  const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
  uint32_t n = 0;
  for (auto &p : c->phases) {
    if (n >= cap) break;
    names_out[n] = p.first;
    ms_out[n] = p.second;
    n++;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------- device-resident flow
static std::vector<fl_t> draw_blinds(uint8_t *tape_state, size_t L) {
  ProverTape t(nullptr, 0, fl_zero());
  memcpy(&t, tape_state, sizeof(t));
  std::vector<fl_t> blinds = t.vector("poly_blinds", L);  // Spartan/src/dense_mlpoly.rs:207-210
  memcpy(tape_state, &t, sizeof(t));
  return blinds;
}
vpin_status vpin_dev_poly_commit(vpin_ctx *ctx, const vpin_gens *gens, const void *d_Z, uint64_t n, uint8_t *tape_state,
                                 void *d_points_out, void *d_blinds_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const SnarkGens *sg = reinterpret_cast<const SnarkGens *>(gens);
  VPIN_REQUIRE(sg && d_Z && d_points_out && d_blinds_out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  VPIN_REQUIRE(n == sg->sat_pc.L * sg->sat_pc.R, VPIN_ERR_SIZE_MISMATCH, "polynomial size does not match gens");
  size_t L = sg->sat_pc.L;
  std::vector<fl_t> blinds = tape_state ? draw_blinds(tape_state, L) : std::vector<fl_t>(L, fl_zero());
  VPIN_CUDA(cudaMemcpyAsync(d_blinds_out, blinds.data(), L * sizeof(fl_t), cudaMemcpyHostToDevice, c_->st));
  hyrax_rows(c_, *sg->sat_label, (const fl_t *)d_Z, L, sg->sat_pc.R, sg->sat_pc.R, tape_state ? (const fl_t *)d_blinds_out : nullptr,
             sg->sat_pc.h_index, nullptr, (uint8_t *)d_points_out);
  c_->sync();  // `blinds` (pageable) must outlive the copy
  VPIN_CATCH
}
vpin_status vpin_dev_poly_commit_with_blinds(vpin_ctx *ctx, const vpin_gens *gens, const void *d_Z, uint64_t n, const void *d_blind1,
                                             const void *d_blind2, void *d_points_out, void *d_blinds_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const SnarkGens *sg = reinterpret_cast<const SnarkGens *>(gens);
  VPIN_REQUIRE(sg && d_Z && d_blind1 && d_blind2 && d_points_out && d_blinds_out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  VPIN_REQUIRE(n == sg->sat_pc.L * sg->sat_pc.R, VPIN_ERR_SIZE_MISMATCH, "polynomial size does not match gens");
  size_t L = sg->sat_pc.L;
  launch_add_vec((const fl_t *)d_blind1, (const fl_t *)d_blind2, L, (fl_t *)d_blinds_out, c_->st);  // commit_test.rs:44-47
  hyrax_rows(c_, *sg->sat_label, (const fl_t *)d_Z, L, sg->sat_pc.R, sg->sat_pc.R, (const fl_t *)d_blinds_out, sg->sat_pc.h_index, nullptr,
             (uint8_t *)d_points_out);
  VPIN_CATCH
}
vpin_status vpin_dev_commitments_add(vpin_ctx *ctx, const void *d_c1, const void *d_c2, uint64_t L, void *d_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(d_c1 && d_c2 && d_out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  DevVec<uint8_t> ok(2 * L, c_->st);
  DevVec<ge_t> p1(L, c_->st), p2(L, c_->st);
  launch_decompress((const uint8_t *)d_c1, L, p1.p, ok.p, c_->st);
  launch_decompress((const uint8_t *)d_c2, L, p2.p, ok.p + L, c_->st);
  launch_points_add(p1.p, p2.p, L, p1.p, c_->st);
  launch_compress(p1.p, L, (uint8_t *)d_out, c_->st);
  std::vector<uint8_t> hok(2 * L);
  ok.download(hok.data(), 2 * L);
  c_->sync();
  for (uint8_t v : hok) VPIN_REQUIRE(v == 1, VPIN_ERR_BAD_ARGUMENT, "point does not decompress");
  VPIN_CATCH
}
vpin_status vpin_witness_from_device(vpin_ctx *ctx, const vpin_gens *gens, const void *d_vars, uint64_t n_vars, const void *d_comm_points,
                                     const void *d_blinds, uint64_t L, vpin_witness **out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const SnarkGens *sg = reinterpret_cast<const SnarkGens *>(gens);
  VPIN_REQUIRE(sg && d_vars && d_comm_points && d_blinds && out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  VPIN_REQUIRE(n_vars == sg->sat_pc.L * sg->sat_pc.R && L == sg->sat_pc.L, VPIN_ERR_SIZE_MISMATCH, "witness size does not match gens");
  auto w = std::make_unique<Witness>();
  w->n_vars = n_vars;
  w->d_vars.alloc(n_vars, c_->st);
  w->d_blinds.alloc(L, c_->st);
  VPIN_CUDA(cudaMemcpyAsync(w->d_vars.p, d_vars, n_vars * sizeof(fl_t), cudaMemcpyDeviceToDevice, c_->st));
  VPIN_CUDA(cudaMemcpyAsync(w->d_blinds.p, d_blinds, L * sizeof(fl_t), cudaMemcpyDeviceToDevice, c_->st));
  w->comm.resize(32 * L);
  VPIN_CUDA(cudaMemcpyAsync(w->comm.data(), d_comm_points, 32 * L, cudaMemcpyDeviceToHost, c_->st));
  c_->sync();
  *out = reinterpret_cast<vpin_witness *>(w.release());
  VPIN_CATCH
}

// ---------------------------------------------------------------------------------------------- kernel-class timing
vpin_status vpin_profile_enable(vpin_ctx *ctx, int32_t on, double min_units) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  prof_drain(c_);
  for (int i = 0; i < PROF_COUNT; i++) c_->prof.acc[i] = Prof::Acc();
  c_->prof.on = on != 0;
  if (min_units >= 0) c_->prof.min_units = min_units;
  c_->d_counters.zero();
  VPIN_CATCH
}
uint32_t vpin_profile_read(vpin_ctx *ctx, const char **names_out, double *ms_out, uint64_t *launches_out, double *units_out,
                           double *bytes_out, uint32_t cap, uint64_t *msm_madds_out) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  uint32_t n = 0;
  try {
    prof_drain(c);
    unsigned long long madds = 0;
    VPIN_CUDA(cudaMemcpyAsync(&madds, c->d_counters.p, sizeof(madds), cudaMemcpyDeviceToHost, c->st));
    c->sync();
    if (msm_madds_out) *msm_madds_out = madds;
  } catch (const std::exception &e) {
    c->err = e.what();
    return 0;
  }
  for (int i = 0; i < PROF_COUNT && n < cap; i++) {
    if (!c->prof.acc[i].launches) continue;
    names_out[n] = prof_class_name(i);
    ms_out[n] = c->prof.acc[i].ms;
    launches_out[n] = c->prof.acc[i].launches;
    units_out[n] = c->prof.acc[i].units;
    bytes_out[n] = c->prof.acc[i].bytes;
    n++;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------- builders
void vpin_point_mult_dims(uint64_t m, uint64_t dims_out[4]) { point_mult_dims(m, dims_out); }
void vpin_point_add_dims(uint64_t n, uint64_t dims_out[4]) { point_add_dims(n, dims_out); }
vpin_status vpin_build_point_mult(vpin_ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32, const uint8_t *py32,
                                  vpin_instance **inst, uint64_t dims_out[4], uint8_t *vars_para32, uint8_t *vars_input32, uint8_t *vars32,
                                  uint8_t *inputs32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(weights_lo_hi && px32 && py32 && inst && dims_out && vars_para32 && vars_input32 && vars32 && inputs32, VPIN_ERR_BAD_ARGUMENT,
               "null argument");
  *inst = reinterpret_cast<vpin_instance *>(build_point_mult(c_, m, weights_lo_hi, px32, py32, dims_out, vars_para32, vars_input32, vars32, inputs32).release());
  VPIN_CATCH
}
vpin_status vpin_build_point_mult_device(vpin_ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32, const uint8_t *py32,
                                         vpin_instance **inst, uint64_t dims_out[4], void *d_vars_para, void *d_vars_input, void *d_vars,
                                         uint64_t padded, uint8_t *inputs32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(weights_lo_hi && px32 && py32 && inst && dims_out && d_vars_para && d_vars_input && d_vars && inputs32, VPIN_ERR_BAD_ARGUMENT,
               "null argument");
  *inst = reinterpret_cast<vpin_instance *>(build_point_mult_dev(c_, m, weights_lo_hi, px32, py32, dims_out, (fl_t *)d_vars_para,
                                                                 (fl_t *)d_vars_input, (fl_t *)d_vars, padded, inputs32).release());
  VPIN_CATCH
}
vpin_status vpin_build_point_add(vpin_ctx *ctx, uint64_t n, const uint8_t *px32, const uint8_t *py32, const uint8_t *rx32,
                                 const uint8_t *ry32, const int64_t *rz_flags, vpin_instance **inst, uint64_t dims_out[4],
                                 uint8_t *vars_para32, uint8_t *vars_input32, uint8_t *vars32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(px32 && py32 && rx32 && ry32 && rz_flags && inst && dims_out && vars_para32 && vars_input32 && vars32, VPIN_ERR_BAD_ARGUMENT,
               "null argument");
  *inst = reinterpret_cast<vpin_instance *>(build_point_add(c_, n, px32, py32, rx32, ry32, rz_flags, dims_out, vars_para32, vars_input32, vars32).release());
  VPIN_CATCH
}

// ---------------------------------------------------------------------------------------------- kernel-level
vpin_status vpin_derive_gens(vpin_ctx *ctx, const char *label, uint64_t n, uint8_t *points_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(label && points_out, VPIN_ERR_BAD_ARGUMENT, "null argument");
  auto g = get_label_gens(c_, label, n + 1);
  DevVec<uint8_t> d(32 * (n + 1), c_->st);
  launch_compress(g->d_pts.p, n + 1, d.p, c_->st);
  d.download(points_out, 32 * (n + 1));
  c_->sync();
  VPIN_CATCH
}
vpin_status vpin_msm(vpin_ctx *ctx, const char *label, const uint8_t *scalars32, uint64_t n, uint8_t out_point[32]) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(label && scalars32 && out_point && n > 0, VPIN_ERR_BAD_ARGUMENT, "null argument");
  auto g = get_label_gens(c_, label, n);
  DevVec<fl_t> ds = upload_scalars(c_, scalars32, n);
  DevVec<uint8_t> d(32, c_->st);
  hyrax_rows(c_, *g, ds.p, 1, n, n, nullptr, 0, nullptr, d.p);
  d.download(out_point, 32);
  c_->sync();
  VPIN_CATCH
}
static void hyrax_dims(uint64_t n, size_t *L, size_t *R) {
  size_t ell = log2_ceil(n);
  *L = (size_t)1 << (ell / 2);
  *R = (size_t)1 << (ell - ell / 2);
}
vpin_status vpin_hyrax_commit(vpin_ctx *ctx, const char *label, const uint8_t *Z32, uint64_t n, const uint8_t *blinds32, uint8_t *points_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(label && Z32 && points_out && is_pow2(n), VPIN_ERR_BAD_ARGUMENT, "bad argument");
  size_t L, R;
  hyrax_dims(n, &L, &R);
  auto g = get_label_gens(c_, label, R + 2);  // DotProductProofGens::new(R): G[0..R), gens_1.G[0] = S[R], h = S[R+1]
  DevVec<fl_t> dZ = upload_scalars(c_, Z32, n);
  DevVec<fl_t> dB;
  if (blinds32) dB = upload_scalars(c_, blinds32, L);
  DevVec<uint8_t> dC(32 * L, c_->st);
  hyrax_rows(c_, *g, dZ.p, L, R, R, blinds32 ? dB.p : nullptr, R + 1, nullptr, dC.p);
  dC.download(points_out, 32 * L);
  c_->sync();
  VPIN_CATCH
}
vpin_status vpin_dev_hyrax_commit(vpin_ctx *ctx, const char *label, const void *d_Z, uint64_t n, const void *d_blinds, void *d_points_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(label && d_Z && d_points_out && is_pow2(n), VPIN_ERR_BAD_ARGUMENT, "bad argument");
  size_t L, R;
  hyrax_dims(n, &L, &R);
  auto g = get_label_gens(c_, label, R + 2);
  hyrax_rows(c_, *g, (const fl_t *)d_Z, L, R, R, (const fl_t *)d_blinds, R + 1, nullptr, (uint8_t *)d_points_out);
  VPIN_CATCH
}
vpin_status vpin_spmv_abc(vpin_ctx *ctx, const vpin_instance *inst, const uint8_t *z32, uint8_t *Az32, uint8_t *Bz32, uint8_t *Cz32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  VPIN_REQUIRE(I && z32 && Az32 && Bz32 && Cz32, VPIN_ERR_BAD_ARGUMENT, "null argument");
  DevVec<fl_t> dz = upload_scalars(c_, z32, 2 * I->num_vars);
  DevVec<fl_t> out(3 * I->num_cons, c_->st);
  for (int k = 0; k < 3; k++) launch_spmv_csr(csr_of(I->M[k], I->num_cons), dz.p, out.p + k * I->num_cons, c_->st);
  uint8_t *dst[3] = {Az32, Bz32, Cz32};
  for (int k = 0; k < 3; k++) download_scalars(c_, out.p + k * I->num_cons, I->num_cons, dst[k]);
  VPIN_CATCH
}
vpin_status vpin_dev_spmv_abc(vpin_ctx *ctx, const vpin_instance *inst, const void *d_z, void *dAz, void *dBz, void *dCz) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  VPIN_REQUIRE(I && d_z && dAz && dBz && dCz, VPIN_ERR_BAD_ARGUMENT, "null argument");
  void *dst[3] = {dAz, dBz, dCz};
  for (int k = 0; k < 3; k++) launch_spmv_csr(csr_of(I->M[k], I->num_cons), (const fl_t *)d_z, (fl_t *)dst[k], c_->st);
  VPIN_CATCH
}
vpin_status vpin_spmv_t_abc(vpin_ctx *ctx, const vpin_instance *inst, const uint8_t *x32, uint8_t *At32, uint8_t *Bt32, uint8_t *Ct32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  const Instance *I = reinterpret_cast<const Instance *>(inst);
  VPIN_REQUIRE(I && x32 && At32 && Bt32 && Ct32, VPIN_ERR_BAD_ARGUMENT, "null argument");
  DevVec<fl_t> dx = upload_scalars(c_, x32, I->num_cons);
  size_t nc = 2 * I->num_vars;
  DevVec<fl_t> out(3 * nc, c_->st), one(1, c_->st);
  fl_t h1 = fl_one();
  one.upload(&h1, 1);
  for (int k = 0; k < 3; k++) launch_spmv_csc_scaled(csc_of(I->M[k], nc), dx.p, one.p, false, out.p + k * nc, c_->d_partials.p, c_->d_partials.n, c_->st);
  uint8_t *dst[3] = {At32, Bt32, Ct32};
  for (int k = 0; k < 3; k++) download_scalars(c_, out.p + k * nc, nc, dst[k]);
  VPIN_CATCH
}
vpin_status vpin_eq_evals(vpin_ctx *ctx, const uint8_t *r32, uint32_t ell, uint8_t *out32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(r32 && out32 && ell <= 30, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  DevVec<fl_t> dr = upload_scalars(c_, r32, ell);
  size_t n = (size_t)1 << ell;
  DevVec<fl_t> out(n, c_->st), tmp(eq_tmp_elems(ell), c_->st);
  launch_eq_evals(dr.p, (int)ell, out.p, tmp.p, c_->st);
  download_scalars(c_, out.p, n, out32);
  VPIN_CATCH
}
vpin_status vpin_dev_eq_evals(vpin_ctx *ctx, const void *d_r, uint32_t ell, void *d_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  DevVec<fl_t> tmp(eq_tmp_elems(ell), c_->st);
  launch_eq_evals((const fl_t *)d_r, (int)ell, (fl_t *)d_out, tmp.p, c_->st);
  VPIN_CATCH
}
vpin_status vpin_sumcheck_cubic_round(vpin_ctx *ctx, const uint8_t *A32, const uint8_t *B32, const uint8_t *C32, const uint8_t *D32,
                                      uint64_t len, uint8_t out96[96]) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(A32 && B32 && C32 && D32 && out96 && is_pow2(len) && len >= 2, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  DevVec<fl_t> a = upload_scalars(c_, A32, len), b = upload_scalars(c_, B32, len), c = upload_scalars(c_, C32, len), d = upload_scalars(c_, D32, len);
  launch_cubic_additive_round(a.p, b.p, c.p, d.p, len / 2, c_->d_small.p, c_->d_partials.p, c_->st);
  download_scalars(c_, c_->d_small.p, 3, out96);
  VPIN_CATCH
}
vpin_status vpin_sumcheck_fused(vpin_ctx *ctx, uint32_t degree, const uint8_t *A32, const uint8_t *B32, const uint8_t *C32, const uint8_t *D32,
                                uint64_t len, const uint8_t *r32, uint8_t *evals_out, uint8_t *finals_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE((degree == 2 || degree == 3) && A32 && B32 && r32 && evals_out && finals_out && is_pow2(len) && len >= 2 &&
                   (degree == 2 || (C32 && D32)), VPIN_ERR_BAD_ARGUMENT, "bad argument");
  cudaStream_t st = c_->st;
  size_t rounds = log2_ceil(len), ntab = degree == 3 ? 4 : 2;
  DevVec<fl_t> t[4];
  const uint8_t *src[4] = {A32, B32, C32, D32};
  for (size_t k = 0; k < ntab; k++) t[k] = upload_scalars(c_, src[k], len);
  std::vector<fl_t> r(rounds), evals(rounds * degree), fin(ntab);
  for (size_t j = 0; j < rounds; j++) VPIN_REQUIRE(fl_from_bytes(r32 + 32 * j, &r[j]), VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
  auto wait = [&](int slot, uint32_t seq) {
    volatile uint32_t *flag = &c_->h_slots[slot].seq;
    while (*flag != seq) {
      cudaError_t e = cudaStreamQuery(st);
      if (e == cudaSuccess && *flag != seq) throw Error(VPIN_ERR_CUDA, "round result never arrived");
      if (e != cudaSuccess && e != cudaErrorNotReady) VPIN_CUDA(e);
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return c_->h_slots[slot].vals;
  };
  auto launch = [&](size_t j, uint32_t *seq) {  // round j: bind with r[j-1] (j > 0), evaluate over q = len >> (j+1) items
    RoundCtl ctl{c_->d_partials.p, c_->d_round_counters.p, c_->d_slots + (j & 1), *seq = ++c_->round_seq};
    size_t q = len >> (j + 1);
    fl_t rp = j ? r[j - 1] : fl_zero();
    if (degree == 3) launch_round_cubic_additive(t[0].p, t[1].p, t[2].p, t[3].p, q, j > 0, rp, ctl, st);
    else launch_round_quad(t[0].p, t[1].p, q, j > 0, rp, ctl, st);
  };
  uint32_t seq;
  launch(0, &seq);
  for (size_t j = 0; j < rounds; j++) {
    memcpy(&evals[j * degree], wait((int)(j & 1), seq), degree * sizeof(fl_t));
    if (j + 1 < rounds) launch(j + 1, &seq);
  }
  FinalArgs fa;
  fa.n = (int)ntab;
  for (size_t k = 0; k < ntab; k++) fa.p[k] = t[k].p;
  RoundCtl ctl{c_->d_partials.p, c_->d_round_counters.p, c_->d_slots + (rounds & 1), seq = ++c_->round_seq};
  launch_round_final(fa, true, r[rounds - 1], ctl, st);
  memcpy(fin.data(), wait((int)(rounds & 1), seq), ntab * sizeof(fl_t));
  for (size_t i = 0; i < evals.size(); i++) fl_to_bytes(evals[i], evals_out + 32 * i);
  for (size_t i = 0; i < ntab; i++) fl_to_bytes(fin[i], finals_out + 32 * i);
  VPIN_CATCH
}
vpin_status vpin_spark_timestamps(vpin_ctx *ctx, const uint32_t *addr_a, uint64_t n_a, const uint32_t *addr_b, uint64_t n_b,
                                  const uint32_t *addr_c, uint64_t n_c, uint64_t N, uint64_t M, uint32_t *addr_out, uint32_t *read_ts_out,
                                  uint32_t *audit_ts_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(addr_out && read_ts_out && audit_ts_out && N && M && n_a <= N && n_b <= N && n_c <= N && 3 * N < ((uint64_t)1 << 32) &&
                   (addr_a || !n_a) && (addr_b || !n_b) && (addr_c || !n_c), VPIN_ERR_BAD_ARGUMENT, "bad argument");
  cudaStream_t st = c_->st;
  const uint32_t *src[3] = {addr_a, addr_b, addr_c};
  size_t nnz[3] = {n_a, n_b, n_c};
  for (int k = 0; k < 3; k++)
    for (size_t i = 0; i < nnz[k]; i++) VPIN_REQUIRE(src[k][i] < M, VPIN_ERR_INVALID_INDEX, "address out of range");
  DevVec<uint32_t> d_in[3], d_addr(3 * N, st), d_ts(3 * N, st), d_audit(M, st), scratch(spark_timestamps_scratch_words(N, M), st);
  const uint32_t *dp[3];
  for (int k = 0; k < 3; k++) {
    d_in[k].alloc(std::max<size_t>(nnz[k], 1), st);
    if (nnz[k]) d_in[k].upload(src[k], nnz[k]);
    dp[k] = d_in[k].p;
  }
  launch_spark_timestamps(dp, nnz, N, M, d_addr.p, d_ts.p, d_audit.p, scratch.p, st);
  d_addr.download(addr_out, 3 * N);
  d_ts.download(read_ts_out, 3 * N);
  d_audit.download(audit_ts_out, M);
  c_->sync();
  VPIN_CATCH
}
vpin_status vpin_dev_cubic_round(vpin_ctx *ctx, const void *dA, const void *dB, const void *dC, const void *dD, uint64_t len, void *d_out3) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  launch_cubic_additive_round((const fl_t *)dA, (const fl_t *)dB, (const fl_t *)dC, (const fl_t *)dD, len / 2, (fl_t *)d_out3, c_->d_partials.p, c_->st);
  VPIN_CATCH
}
vpin_status vpin_sumcheck_quad_round(vpin_ctx *ctx, const uint8_t *A32, const uint8_t *B32, uint64_t len, uint8_t out64[64]) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(A32 && B32 && out64 && is_pow2(len) && len >= 2, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  DevVec<fl_t> a = upload_scalars(c_, A32, len), b = upload_scalars(c_, B32, len);
  launch_quad_round(a.p, b.p, len / 2, c_->d_small.p, c_->d_partials.p, c_->st);
  download_scalars(c_, c_->d_small.p, 2, out64);
  VPIN_CATCH
}
vpin_status vpin_dev_quad_round(vpin_ctx *ctx, const void *dA, const void *dB, uint64_t len, void *d_out2) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  launch_quad_round((const fl_t *)dA, (const fl_t *)dB, len / 2, (fl_t *)d_out2, c_->d_partials.p, c_->st);
  VPIN_CATCH
}
vpin_status vpin_sumcheck_cubic3_round(vpin_ctx *ctx, const uint8_t *A32, const uint8_t *B32, const uint8_t *C32, uint64_t len,
                                       uint8_t out96[96]) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(A32 && B32 && C32 && out96 && is_pow2(len) && len >= 2, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  DevVec<fl_t> a = upload_scalars(c_, A32, len), b = upload_scalars(c_, B32, len), c = upload_scalars(c_, C32, len);
  const fl_t *hp[3] = {a.p, b.p, c.p};
  DevVec<const fl_t *> dp(3, c_->st);
  dp.upload(hp, 3);
  launch_cubic_batched_round(dp.p, dp.p + 1, dp.p + 2, 1, len / 2, c_->d_small.p, c_->d_partials.p, c_->st);
  download_scalars(c_, c_->d_small.p, 3, out96);
  VPIN_CATCH
}
vpin_status vpin_bind_top(vpin_ctx *ctx, uint8_t *Z32, uint64_t len, const uint8_t r32[32]) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(Z32 && r32 && is_pow2(len) && len >= 2, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  DevVec<fl_t> z = upload_scalars(c_, Z32, len), r = upload_scalars(c_, r32, 1);
  launch_bind_top(z.p, len / 2, r.p, c_->st);
  download_scalars(c_, z.p, len / 2, Z32);
  VPIN_CATCH
}
vpin_status vpin_dev_bind_top(vpin_ctx *ctx, void *dZ, uint64_t len, const void *d_r) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  launch_bind_top((fl_t *)dZ, len / 2, (const fl_t *)d_r, c_->st);
  VPIN_CATCH
}
vpin_status vpin_bound(vpin_ctx *ctx, const uint8_t *Z32, uint64_t len, const uint8_t *L32, uint8_t *out32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(Z32 && L32 && out32 && is_pow2(len), VPIN_ERR_BAD_ARGUMENT, "bad argument");
  size_t L, R;
  hyrax_dims(len, &L, &R);
  DevVec<fl_t> z = upload_scalars(c_, Z32, len), l = upload_scalars(c_, L32, L);
  DevVec<fl_t> out(R, c_->st), tmp(64 * R, c_->st);
  launch_bound(z.p, l.p, L, R, out.p, tmp.p, c_->st);
  download_scalars(c_, out.p, R, out32);
  VPIN_CATCH
}
vpin_status vpin_product_tree(vpin_ctx *ctx, const uint8_t *leaves32, uint64_t n, uint8_t *tree32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(leaves32 && tree32 && is_pow2(n) && n >= 2, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  DevVec<fl_t> tree(2 * n, c_->st);
  tree.zero();
  {
    DevVec<fl_t> leaves = upload_scalars(c_, leaves32, n);
    VPIN_CUDA(cudaMemcpyAsync(tree.p, leaves.p, n * sizeof(fl_t), cudaMemcpyDeviceToDevice, c_->st));
    c_->sync();
  }
  build_tree(c_, tree.p, n, c_->st);
  download_scalars(c_, tree.p, 2 * n - 2, tree32);
  VPIN_CATCH
}
vpin_status vpin_hash_layer(vpin_ctx *ctx, const uint32_t *addr, const uint8_t *val32, const uint32_t *ts, uint64_t n,
                            const uint8_t gamma32[32], const uint8_t tau32[32], uint8_t *read32, uint8_t *write32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(addr && val32 && ts && gamma32 && tau32 && read32 && write32 && n, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  DevVec<fl_t> val = upload_scalars(c_, val32, n);
  uint8_t gt[64];
  memcpy(gt, gamma32, 32);
  memcpy(gt + 32, tau32, 32);
  DevVec<fl_t> d_gt = upload_scalars(c_, gt, 2);
  DevVec<uint32_t> d_addr(n, c_->st), d_ts(n, c_->st);
  d_addr.upload(addr, n);
  d_ts.upload(ts, n);
  DevVec<fl_t> rd(n, c_->st), wr(n, c_->st);
  launch_hash_ops(d_addr.p, val.p, d_ts.p, n, d_gt.p, rd.p, wr.p, c_->st);
  download_scalars(c_, rd.p, n, read32);
  download_scalars(c_, wr.p, n, write32);
  VPIN_CATCH
}
vpin_status vpin_deref_gather(vpin_ctx *ctx, const uint32_t *addr, uint64_t n, const uint8_t *mem32, uint64_t num_cells, uint8_t *out32) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(addr && mem32 && out32 && n && num_cells, VPIN_ERR_BAD_ARGUMENT, "bad argument");
  for (uint64_t i = 0; i < n; i++) VPIN_REQUIRE(addr[i] < num_cells, VPIN_ERR_INVALID_INDEX, "address out of range");
  DevVec<fl_t> mem = upload_scalars(c_, mem32, num_cells);
  DevVec<uint32_t> d_addr(n, c_->st);
  d_addr.upload(addr, n);
  DevVec<fl_t> out(n, c_->st);
  launch_gather(d_addr.p, mem.p, n, out.p, c_->st);
  download_scalars(c_, out.p, n, out32);
  VPIN_CATCH
}
vpin_status vpin_dev_to_mont(vpin_ctx *ctx, const void *d_in, uint64_t n, void *d_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  launch_to_mont((const fl_t *)d_in, n, (fl_t *)d_out, c_->st);
  VPIN_CATCH
}
vpin_status vpin_dev_from_mont(vpin_ctx *ctx, const void *d_in, uint64_t n, void *d_out) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  launch_from_mont((const fl_t *)d_in, n, (fl_t *)d_out, c_->st);
  VPIN_CATCH
}
vpin_status vpin_imad_peak(vpin_ctx *ctx, double *macs_per_second) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(macs_per_second, VPIN_ERR_BAD_ARGUMENT, "null argument");
  *macs_per_second = measure_imad_peak(c_);
  VPIN_CATCH
}
vpin_status vpin_imad_peak_forms(vpin_ctx *ctx, double forms[2]) {
  VPIN_TRY(reinterpret_cast<Ctx *>(ctx))
  VPIN_REQUIRE(forms, VPIN_ERR_BAD_ARGUMENT, "null argument");
  measure_imad_peaks(c_, forms);
  VPIN_CATCH
}

}  // extern "C"
