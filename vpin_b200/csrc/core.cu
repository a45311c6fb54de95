#include "core.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <time.h>

#include <algorithm>
#include <mutex>

namespace vpin {

LaunchCounter g_kernel_launches;

// ------------------------------------------------------------------------------------------------ block cache
struct BlockCache {
  std::mutex mu;
  std::map<size_t, std::vector<void *>> free_;   // rounded size -> idle blocks
  std::map<void *, size_t> live_;                // every block handed out by this cache -> rounded size
  std::function<bool()> pressure;                // see block_cache_set_pressure_hook
};
namespace {
std::mutex g_cache_mu;
std::map<cudaStream_t, std::unique_ptr<BlockCache>> g_caches;
std::map<void *, size_t> g_orphans;  // blocks whose cache died while they were in use (freed on release)
size_t round_block(size_t bytes) {
  if (bytes <= 256) return 256;
  if (bytes >= ((size_t)2 << 20)) return (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  size_t r = 256;
  while (r < bytes) r <<= 1;
  return r;
}
}  // namespace
std::shared_mutex &device_sync_gate() {
  static std::shared_mutex mu;
  return mu;
}
void gated_cuda_free(void *p) {
  std::unique_lock<std::shared_mutex> g(device_sync_gate());
  cudaFree(p);
}
void gated_cuda_free_host(void *p) {
  std::unique_lock<std::shared_mutex> g(device_sync_gate());
  cudaFreeHost(p);
}
BlockCache *block_cache_of(cudaStream_t st) {
  std::lock_guard<std::mutex> g(g_cache_mu);
  auto it = g_caches.find(st);
  return it == g_caches.end() ? nullptr : it->second.get();
}
void block_cache_register(cudaStream_t st) {
  std::lock_guard<std::mutex> g(g_cache_mu);
  g_caches[st] = std::make_unique<BlockCache>();
}
void block_cache_set_pressure_hook(cudaStream_t st, std::function<bool()> hook) {
  std::lock_guard<std::mutex> g(g_cache_mu);
  auto it = g_caches.find(st);
  if (it != g_caches.end()) it->second->pressure = std::move(hook);
}
void block_cache_unregister(cudaStream_t st) {
  std::unique_ptr<BlockCache> c;
  {
    std::lock_guard<std::mutex> g(g_cache_mu);
    auto it = g_caches.find(st);
    if (it == g_caches.end()) return;
    c = std::move(it->second);
    g_caches.erase(it);
    for (auto &kv : c->live_) g_orphans.insert(kv);  // still owned by some handle: freed when it lets go
  }
  for (auto &kv : c->free_)
    for (void *p : kv.second) gated_cuda_free(p);
}
void *block_cache_alloc(BlockCache *c, size_t bytes) {
  size_t r = round_block(bytes);
  void *p = nullptr;
  if (c) {
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->free_.find(r);
    if (it != c->free_.end() && !it->second.empty()) {
      p = it->second.back();
      it->second.pop_back();
      c->live_[p] = r;
      return p;
    }
  }
  cudaError_t e = cudaMalloc(&p, r);
  // Out of memory: give idle blocks back to the driver, largest first and only until the request fits (returning everything
  // makes the next proof re-allocate it all: at LeNet layer 5, where the cache holds ~170 of the 180 GB, that thrash doubled
  // the time of whole phases); then ask the owner to let go of what it can spare (the idle SPARK workspace of a previous
  // proof) and repeat.
  for (int attempt = 0; e != cudaSuccess && c && attempt < 2; attempt++) {
    cudaGetLastError();
    if (attempt == 1 && !(c->pressure && c->pressure())) break;
    for (;;) {
      size_t freed = 0;
      std::vector<void *> give_back;  // (freed outside the cache lock: the free waits for the device-synchronisation gate)
      {
        std::lock_guard<std::mutex> g(c->mu);
        for (auto it = c->free_.rbegin(); it != c->free_.rend() && freed < r; ++it) {
          while (!it->second.empty() && freed < r) {
            give_back.push_back(it->second.back());
            it->second.pop_back();
            freed += it->first;
          }
        }
        for (auto it = c->free_.begin(); it != c->free_.end();) it = it->second.empty() ? c->free_.erase(it) : std::next(it);
      }
      for (void *q : give_back) gated_cuda_free(q);
      e = cudaMalloc(&p, r);
      if (e == cudaSuccess || freed == 0) break;  // done, or nothing left to give back
      cudaGetLastError();
    }
  }
  if (e != cudaSuccess) {  // last resort: the idle blocks of every other context of the process (concurrent proofs share the GPU)
    cudaGetLastError();
    std::vector<BlockCache *> others;
    std::vector<void *> give_back;
    {
      std::lock_guard<std::mutex> g(g_cache_mu);
      for (auto &kv : g_caches)
        if (kv.second.get() != c) others.push_back(kv.second.get());
      for (BlockCache *o : others) {  // (under g_cache_mu: a cache cannot be unregistered meanwhile)
        std::lock_guard<std::mutex> go(o->mu);
        for (auto &kv : o->free_)
          for (void *q : kv.second) give_back.push_back(q);
        o->free_.clear();
      }
    }
    for (void *q : give_back) gated_cuda_free(q);
    e = cudaMalloc(&p, r);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(VPIN_ERR_OOM, std::string("cudaMalloc(") + std::to_string(r) + "): " + cudaGetErrorString(e));
  }
  if (c) {
    std::lock_guard<std::mutex> g(c->mu);
    c->live_[p] = r;
  } else {
    std::lock_guard<std::mutex> g(g_cache_mu);
    g_orphans[p] = r;
  }
  return p;
}
void block_cache_free(BlockCache *c, void *p) {
  if (c) {
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->live_.find(p);
    if (it != c->live_.end()) {
      c->free_[it->second].push_back(p);
      c->live_.erase(it);
      return;
    }
  }
  {
    std::lock_guard<std::mutex> g(g_cache_mu);
    g_orphans.erase(p);
  }
  gated_cuda_free(p);  // implicit device synchronisation: nothing can still be using it
}

// ------------------------------------------------------------------------------------------------ profiling
const char *prof_class_name(int cls) {
  static const char *names[PROF_COUNT] = {"msm_recode", "msm_accumulate", "msm_finish", "sumcheck_cubic_round", "sumcheck_quad_round",
                                          "sumcheck_batched_round", "bind_top", "spmv_csr", "spmv_csc", "eq_evals", "product_tree",
                                          "hash_layer", "deref_gather", "bound_LZ", "dot", "bullet_round", "round_final"};
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

// resident MSM blocks per SM while a commitment runs on the side stream (VPIN_SIDE_MSM_BLOCKS, default 2 of the usual 6: with the
// single block of a background context's MSM that still leaves the registers of one 256-thread round-kernel block on every SM)
int side_msm_blocks_per_sm() {
  static const int v = [] { const char *e = getenv("VPIN_SIDE_MSM_BLOCKS"); int x = e ? atoi(e) : 2; return x < 0 ? 0 : x; }();
  return v;
}
SideScope::SideScope(Ctx *ctx) : c(ctx) {
  VPIN_REQUIRE(!c->on_side && c->world == 1, VPIN_ERR_BAD_ARGUMENT, "side stream: nested or distributed use");
  if (!c->st_side) {
    // the LOWEST priority of the device: what runs here fills gaps, the main stream's round kernels take free resources first
    int lo = 0, hi = 0;
    VPIN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    VPIN_CUDA(cudaStreamCreateWithPriority(&c->st_side, cudaStreamNonBlocking, lo));
    block_cache_register(c->st_side);
    VPIN_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    VPIN_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  VPIN_CUDA(cudaEventRecord(c->ev_fork, c->st));
  VPIN_CUDA(cudaStreamWaitEvent(c->st_side, c->ev_fork, 0));
  c->st_main = c->st;
  c->st = c->st_side;
  c->on_side = true;
}
SideScope::~SideScope() {
  cudaEventRecord(c->ev_join, c->st_side);
  c->st = c->st_main;
  c->on_side = false;
}
void SideScope::join(Ctx *ctx) { VPIN_CUDA(cudaStreamWaitEvent(ctx->st, ctx->ev_join, 0)); }

// ------------------------------------------------------------------------------------------------ multi-GPU plumbing
// NCCL is resolved at run time (dlopen) so that the library has no link-time dependency on it and a process that has
// already loaded torch's bundled libnccl.so.2 shares that copy.
namespace {
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &nccl() {
  static NcclApi api;
  if (api.h) return api;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  VPIN_REQUIRE(h, VPIN_ERR_CUDA, std::string("cannot load libnccl.so.2: ") + dlerror());
#define VPIN_NCCL_SYM(field, name)                                       \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));     \
  VPIN_REQUIRE(api.field, VPIN_ERR_CUDA, "libnccl is missing " name)
  VPIN_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  VPIN_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  VPIN_NCCL_SYM(AllGather, "ncclAllGather");
  VPIN_NCCL_SYM(AllReduce, "ncclAllReduce");
  VPIN_NCCL_SYM(Broadcast, "ncclBroadcast");
  VPIN_NCCL_SYM(GroupStart, "ncclGroupStart");
  VPIN_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  VPIN_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  VPIN_NCCL_SYM(CommAbort, "ncclCommAbort");
  VPIN_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef VPIN_NCCL_SYM
  api.h = h;
  return api;
}
#define VPIN_NCCL(x)                                                                                             \
  do {                                                                                                           \
    ncclResult_t r_ = (x);                                                                                       \
    if (r_ != ncclSuccess) throw ::vpin::Error(VPIN_ERR_CUDA, std::string(#x) + ": " + nccl().GetErrorString(r_)); \
  } while (0)
}  // namespace

bool shard_rows(size_t rows, int rank, int world, size_t *r0, size_t *r1) {
  *r0 = 0;
  *r1 = rows;
  if (world <= 1 || rows % (size_t)world != 0 || rows / (size_t)world < kMinShardRows) return false;
  size_t per = rows / (size_t)world;
  *r0 = per * (size_t)rank;
  *r1 = *r0 + per;
  return true;
}
void dist_get_unique_id(uint8_t out[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  VPIN_NCCL(nccl().GetUniqueId(&id));
  memcpy(out, &id, 128);
}
void dist_init(Ctx *ctx, int rank, int world, const uint8_t id_bytes[128]) {
  VPIN_REQUIRE(world >= 1 && rank >= 0 && rank < world, VPIN_ERR_BAD_ARGUMENT, "bad rank / world");
  dist_destroy(ctx);
  ctx->rank = rank;
  ctx->world = world;
  if (world == 1) return;
  ncclUniqueId id;
  memcpy(&id, id_bytes, 128);
  ncclComm_t comm;
  VPIN_CUDA(cudaSetDevice(ctx->device));
  VPIN_NCCL(nccl().CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
}
void dist_destroy(Ctx *ctx) {
  if (ctx->nccl_comm) {
    cudaStreamSynchronize(ctx->st);
    nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  ctx->rank = 0;
  ctx->world = 1;
}
// A peer that failed (out of memory, a prover-side assertion) never joins the next collective: the NCCL kernel of the surviving
// ranks then spins forever and their stream never drains. Called when a round result is overdue on a distributed context: tears
// the communicator down (which releases the stuck kernel) so that the call can fail with an error instead of hanging.
void dist_abort(Ctx *ctx) {
  if (ctx->nccl_comm) {
    nccl().CommAbort((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
  }
}
// in-place broadcasts of several buffers, each from its own root rank, fused into one NCCL group
void dist_broadcast_many(Ctx *ctx, void *const *bufs, const size_t *bytes, const int *roots, int count) {
  VPIN_REQUIRE(ctx->nccl_comm, VPIN_ERR_BAD_ARGUMENT, "context is not distributed");
  VPIN_NCCL(nccl().GroupStart());
  for (int i = 0; i < count; i++)
    VPIN_NCCL(nccl().Broadcast(bufs[i], bufs[i], bytes[i], ncclUint8, roots[i], (ncclComm_t)ctx->nccl_comm, ctx->st));
  VPIN_NCCL(nccl().GroupEnd());
}
void vpin_ctx_impl::sync_distributed() {
  static const double limit_s = [] { const char *e = getenv("VPIN_DIST_TIMEOUT_S"); double v = e ? atof(e) : 120.0; return v > 0 ? v : 120.0; }();
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (uint64_t spins = 1;; spins++) {
    cudaError_t e = cudaStreamQuery(st);
    if (e == cudaSuccess) return;
    if (e != cudaErrorNotReady) VPIN_CUDA(e);
    if ((spins & 0xffff) == 0) {
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > limit_s) {
        dist_abort(this);
        throw Error(VPIN_ERR_CUDA, "stream did not drain on a distributed context (a peer rank failed?): communicator aborted");
      }
    }
    __builtin_ia32_pause();
  }
}
void dist_allgather_inplace(Ctx *ctx, void *buf, size_t bytes_per_rank) {
  VPIN_REQUIRE(ctx->nccl_comm, VPIN_ERR_BAD_ARGUMENT, "context is not distributed");
  const uint8_t *send = (const uint8_t *)buf + (size_t)ctx->rank * bytes_per_rank;
  VPIN_NCCL(nccl().AllGather(send, buf, bytes_per_rank, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->st));
}

// ------------------------------------------------------------------------------------------------ generators
namespace {
std::mutex g_gens_mu;  // held while a table is being built, so that concurrent contexts wait for it instead of building their own
std::map<std::pair<int, std::string>, std::weak_ptr<LabelGens>> g_gens;
}  // namespace
// the generators of a label if this context (or another context of the process on the same device) already holds at least n
std::shared_ptr<LabelGens> find_label_gens(Ctx *ctx, const std::string &label, size_t n) {
  auto it = ctx->label_gens.find(label);
  if (it != ctx->label_gens.end() && it->second->n >= n) return it->second;
  std::lock_guard<std::mutex> global_lock(g_gens_mu);
  auto gi = g_gens.find({ctx->device, label});
  if (gi != g_gens.end())
    if (auto shared = gi->second.lock())
      if (shared->n >= n) return ctx->label_gens[label] = shared;
  return nullptr;
}
std::shared_ptr<LabelGens> get_label_gens(Ctx *ctx, const std::string &label, size_t n, size_t table_budget, bool *built) {
  if (built) *built = false;
  if (auto have = find_label_gens(ctx, label, n)) return have;
  std::lock_guard<std::mutex> global_lock(g_gens_mu);
  {
    auto gi = g_gens.find({ctx->device, label});
    if (gi != g_gens.end())
      if (auto shared = gi->second.lock())
        if (shared->n >= n) return ctx->label_gens[label] = shared;
  }
  if (built) *built = true;
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
  // window width: the widest whose table fits the caller's budget (snark_gens_create plans both tables of a shape against
  // the proof's working set) or, without one, 30 % of the free HBM; at most 64 GiB either way. VPIN_MSM_W pins it.
  {
    size_t free_b = 0, total_b = 0;
    VPIN_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = std::min<size_t>(table_budget ? std::min(table_budget, free_b) : free_b / 10 * 3, (size_t)64 << 30);
    // fewest windows (= mixed additions per scalar) whose table fits; at equal windows four sub-tables (the shorter Horner pass)
    int W = kMsmMinW, sub = kMsmSub;
    for (int w = kMsmMinW; w <= kMsmMaxW; w++)
      for (int sb = 2; sb <= 4; sb += 2) {
        if (msm_table_bytes_per_base(w, sb) * n > budget) continue;
        int have = msm_geom(W, sub).windows, cand = msm_geom(w, sb).windows;
        if (cand < have || (cand == have && sb > sub)) { W = w; sub = sb; }
      }
    if (const char *e = getenv("VPIN_MSM_W")) {  // VPIN_MSM_W / VPIN_MSM_SUB pin the geometry (tests, experiments)
      int w = atoi(e);
      if (w >= kMsmMinW && w <= kMsmMaxW) W = w;
    }
    if (const char *e = getenv("VPIN_MSM_SUB")) {
      int sb = atoi(e);
      if (sb == 2 || sb == 4) sub = sb;
    }
    g->geom = msm_geom(W, sub);
  }
  g->d_table.alloc(msm_table_entries(n, g->geom), ctx->st);
  {
    DevVec<ge_t> scratch(n, ctx->st);
    launch_table_build(g->d_pts.p, n, g->geom, g->d_table.p, scratch.p, ctx->st);
  }
  ctx->sync();
  ctx->label_gens[label] = g;
  g_gens[{ctx->device, label}] = g;
  return g;
}

// ------------------------------------------------------------------------------------------------ Hyrax rows
static void hyrax_rows_local(Ctx *ctx, const LabelGens &g, const fl_t *dZ, size_t rows, size_t cols, size_t ld, const fl_t *d_blinds,
                             size_t blind_base, ge_t *d_points, uint8_t *d_comp) {
  size_t cols_total = cols + (d_blinds ? 1 : 0);
  size_t stride = msm_col_stride(cols_total);
  // bound the digit buffer (2 bytes x windows per scalar) to ~1 GiB per pass
  const MsmGeom &geom = g.geom;
  size_t max_rows = ((size_t)1 << 30) / (stride * geom.windows * sizeof(uint16_t));
  if (max_rows < 1) max_rows = 1;
  size_t chunk = rows < max_rows ? rows : max_rows;
  size_t segs = msm_num_segments(chunk, cols_total, geom);
  DevVec<uint16_t> digits(msm_digits_count(chunk, cols_total, geom), ctx->st);
  DevVec<ge_t> partial(chunk * geom.group * segs, ctx->st), sums(segs > 1 ? chunk * geom.group : 0, ctx->st);
  uint32_t *d_wmask = reinterpret_cast<uint32_t *>(ctx->d_counters.p + (ctx->on_side ? 3 : 2));  // windows in use, per recode launch
  for (size_t r0 = 0; r0 < rows; r0 += chunk) {
    size_t nr = rows - r0 < chunk ? rows - r0 : chunk;
    double pts = (double)nr * cols_total;
    {
      ProfScope ps(ctx, PROF_MSM_RECODE, pts, pts * (32 + 2 * geom.windows));
      launch_recode(dZ + r0 * ld, nr, cols, ld, d_blinds ? d_blinds + r0 : nullptr, geom, digits.p, ctx->d_counters.p, ctx->st, d_wmask);
    }
    {
      ProfScope ps(ctx, PROF_MSM_ACCUMULATE, pts, 0);
      launch_msm_accumulate(g.table(), digits.p, nr, cols, d_blinds != nullptr, blind_base, segs, partial.p, ctx->st, d_wmask,
                            ctx->background ? 1 : (ctx->on_side ? side_msm_blocks_per_sm() : 0));
    }
    ProfScope ps(ctx, PROF_MSM_FINISH, pts, 0);
    launch_msm_finish(partial.p, nr, segs, geom, sums.p, d_points ? d_points + r0 : nullptr, d_comp ? d_comp + 32 * r0 : nullptr, ctx->st);
  }
}
// Multi-GPU: the L rows of a commitment are independent MSMs over replicated generator tables, so rank r commits to
// rows [r L/G, (r+1) L/G) and the compressed rows (32 B each) are exchanged with one in-place NCCL all-gather.
void hyrax_rows(Ctx *ctx, const LabelGens &g, const fl_t *dZ, size_t rows, size_t cols, size_t ld, const fl_t *d_blinds,
                size_t blind_base, ge_t *d_points, uint8_t *d_comp) {
  VPIN_REQUIRE(cols <= g.n && (!d_blinds || blind_base < g.n), VPIN_ERR_SIZE_MISMATCH, "hyrax_rows: not enough generators");
  size_t r0, r1;
  bool sharded = shard_rows(rows, ctx->rank, ctx->world, &r0, &r1) && ctx->nccl_comm;
  if (!sharded) { r0 = 0; r1 = rows; }
  hyrax_rows_local(ctx, g, dZ + r0 * ld, r1 - r0, cols, ld, d_blinds ? d_blinds + r0 : nullptr, blind_base,
                   d_points ? d_points + r0 : nullptr, d_comp ? d_comp + 32 * r0 : nullptr);
  if (sharded) {
    if (d_comp) dist_allgather_inplace(ctx, d_comp, 32 * (r1 - r0));
    if (d_points) dist_allgather_inplace(ctx, d_points, sizeof(ge_t) * (r1 - r0));
  }
}

// ------------------------------------------------------------------------------------------------ instance
// Instance::new on the device: the raw 48-byte COO triples go up as they are; one kernel validates them (first failing
// entry wins, like the reference's sequential loop), converts the values to Montgomery form, applies the column shift of
// Spartan/src/lib.rs:196-200 and counts rows / columns; a prefix sum and a scatter kernel then produce the CSR and CSC
// forms (entry order inside a row or column is irrelevant: field addition is exact and commutative).
namespace {
__global__ void __launch_bounds__(256) k_coo_unpack(const vpin_coo_entry *raw, size_t n, uint64_t num_cons, uint64_t num_vars,
                                                    uint64_t num_cols_in, uint64_t shift, uint32_t *row, uint32_t *col, fl_t *val,
                                                    uint32_t *row_cnt, uint32_t *col_cnt, unsigned long long *first_err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 *q = reinterpret_cast<const uint4 *>(raw + i);
  uint4 rc = __ldg(q), v0 = __ldg(q + 1), v1 = __ldg(q + 2);
  uint64_t r = (uint64_t)rc.x | ((uint64_t)rc.y << 32), c = (uint64_t)rc.z | ((uint64_t)rc.w << 32);
  fl_t x;
  x.v[0] = v0.x; x.v[1] = v0.y; x.v[2] = v0.z; x.v[3] = v0.w; x.v[4] = v1.x; x.v[5] = v1.y; x.v[6] = v1.z; x.v[7] = v1.w;
  int64_t br = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { int64_t d = (int64_t)x.v[k] - (int64_t)fl_modulus_limb(k) + br; br = d >> 32; }
  uint32_t kind = (r >= num_cons || c >= num_cols_in) ? 2u : (br == 0 ? 1u : 0u);  // 2 = InvalidIndex, 1 = InvalidScalar
  if (kind) {
    atomicMin(first_err, ((unsigned long long)i << 2) | kind);
    return;
  }
  uint32_t cc = (uint32_t)(c >= num_vars ? c + shift : c);
  row[i] = (uint32_t)r;
  col[i] = cc;
  fl_t m = fl_to_mont(x);
  uint4 *o = reinterpret_cast<uint4 *>(val + i);
  o[0] = make_uint4(m.v[0], m.v[1], m.v[2], m.v[3]);
  o[1] = make_uint4(m.v[4], m.v[5], m.v[6], m.v[7]);
  atomicAdd(row_cnt + r, 1u);
  atomicAdd(col_cnt + cc, 1u);
}
// Montgomery forms of the dictionary values, code k at index k - 1 (kernels_poly.cuh SpmvCode)
struct CodeBook { fl_t v[5]; };
__device__ __forceinline__ uint8_t spmv_code_of(const uint4 &a, const uint4 &b, const CodeBook &cb) {
  uint8_t code = kCodeGeneral;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const uint32_t *w = cb.v[k].v;
    bool eq = a.x == w[0] && a.y == w[1] && a.z == w[2] && a.w == w[3] && b.x == w[4] && b.y == w[5] && b.z == w[6] && b.w == w[7];
    code = eq ? (uint8_t)(k + 1) : code;
  }
  return code;
}
__global__ void __launch_bounds__(256) k_coo_scatter(const uint32_t *row, const uint32_t *col, const fl_t *val, size_t n, uint32_t *row_cur,
                                                     uint32_t *col_cur, uint32_t *csr_col, fl_t *csr_val, uint32_t *csc_row, fl_t *csc_val,
                                                     CodeBook cb, uint8_t *csr_code, uint8_t *csc_code) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r = row[i], c = col[i];
  const uint4 *q = reinterpret_cast<const uint4 *>(val + i);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  const uint8_t code = spmv_code_of(a, b, cb);
  uint32_t p = atomicAdd(row_cur + r, 1u);
  csr_code[p] = code;
  csr_col[p] = c;
  reinterpret_cast<uint4 *>(csr_val + p)[0] = a;
  reinterpret_cast<uint4 *>(csr_val + p)[1] = b;
  uint32_t p2 = atomicAdd(col_cur + c, 1u);
  csc_code[p2] = code;
  csc_row[p2] = r;
  reinterpret_cast<uint4 *>(csc_val + p2)[0] = a;
  reinterpret_cast<uint4 *>(csc_val + p2)[1] = b;
}
__global__ void __launch_bounds__(256) k_find_long_cols(const uint32_t *cptr, size_t ncols, uint32_t *long_cols, uint32_t *n_long, uint32_t cap,
                                                        uint32_t threshold) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  if (cptr[c + 1] - cptr[c] > threshold) {
    uint32_t k = atomicAdd(n_long, 1u);
    if (k < cap) long_cols[k] = (uint32_t)c;
  }
}

// returns 0, or the error kind of the first invalid entry
uint32_t build_matrix(Ctx *ctx, MatrixDev &m, const vpin_coo_entry *entries, bool entries_on_device, size_t n, size_t n_pad, uint64_t num_cons,
                      uint64_t num_vars, uint64_t num_inputs, size_t num_rows, size_t num_vars_padded) {
  cudaStream_t st = ctx->st;
  size_t num_cols = 2 * num_vars_padded, nnz = n + n_pad;
  m.nnz = nnz;
  uint64_t shift = num_vars_padded - num_vars, ncols_in = num_vars + 1 + num_inputs;
  std::vector<vpin_coo_entry> pad(n_pad);
  for (size_t i = 0; i < n_pad; i++) {  // Spartan/src/lib.rs:207-211: (i, num_vars, 0), column NOT shifted
    memset(&pad[i], 0, sizeof(vpin_coo_entry));
    pad[i].row = n + i;
    pad[i].col = num_vars;
  }
  DevVec<vpin_coo_entry> raw(entries_on_device ? 0 : nnz, st);
  if (n && !entries_on_device) VPIN_CUDA(cudaMemcpyAsync(raw.p, entries, n * sizeof(vpin_coo_entry), cudaMemcpyHostToDevice, st));
  const vpin_coo_entry *d_raw = entries_on_device ? entries : raw.p;
  m.coo_row.alloc(nnz, st); m.coo_col.alloc(nnz, st); m.coo_val.alloc(nnz, st);
  m.csr_ptr.alloc(num_rows + 1, st); m.csc_ptr.alloc(num_cols + 1, st);
  m.csr_col.alloc(nnz, st); m.csr_val.alloc(nnz, st); m.csc_row.alloc(nnz, st); m.csc_val.alloc(nnz, st);
  m.csr_code.alloc(std::max<size_t>(nnz, 1), st); m.csc_code.alloc(std::max<size_t>(nnz, 1), st);
  m.csr_ptr.zero(); m.csc_ptr.zero();
  DevVec<unsigned long long> d_err(1, st);
  VPIN_CUDA(cudaMemsetAsync(d_err.p, 0xff, sizeof(unsigned long long), st));
  if (n)
    ++g_kernel_launches, k_coo_unpack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_raw, n, num_cons, num_vars, ncols_in, shift, m.coo_row.p, m.coo_col.p,
                                                                                 m.coo_val.p, m.csr_ptr.p, m.csc_ptr.p, d_err.p);
  if (n_pad) {  // padding rows carry the unshifted column num_vars and are exempt from the index checks
    DevVec<vpin_coo_entry> dpad(n_pad, st);
    dpad.upload(pad.data(), n_pad);
    ++g_kernel_launches, k_coo_unpack<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(dpad.p, n_pad, num_rows, num_cols, num_cols, 0, m.coo_row.p + n,
                                                                                     m.coo_col.p + n, m.coo_val.p + n, m.csr_ptr.p, m.csc_ptr.p, d_err.p);
    ctx->sync();
  }
  unsigned long long err = 0;
  VPIN_CUDA(cudaMemcpyAsync(&err, d_err.p, sizeof(err), cudaMemcpyDeviceToHost, st));
  ctx->sync();
  if (err != ~0ull) return (uint32_t)(err & 3);
  {  // exclusive prefix sums of the n + 1 counters (cnt[n] == 0 on entry) -> CSR / CSC pointers
    DevVec<uint32_t> scratch(exclusive_scan_scratch_words(std::max(num_rows, num_cols) + 1), st);
    launch_exclusive_scan_u32(m.csr_ptr.p, num_rows + 1, scratch.p, st);
    launch_exclusive_scan_u32(m.csc_ptr.p, num_cols + 1, scratch.p, st);
  }
  DevVec<uint32_t> row_cur(num_rows + 1, st), col_cur(num_cols + 1, st), n_long(1, st);
  VPIN_CUDA(cudaMemcpyAsync(row_cur.p, m.csr_ptr.p, (num_rows + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  VPIN_CUDA(cudaMemcpyAsync(col_cur.p, m.csc_ptr.p, (num_cols + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  CodeBook cb;
  {
    const fl_t one = fl_one(), two = fl_add(one, one);
    cb.v[0] = one; cb.v[1] = fl_neg(one); cb.v[2] = two; cb.v[3] = fl_neg(two); cb.v[4] = fl_add(two, one);
  }
  if (nnz)
    ++g_kernel_launches, k_coo_scatter<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(m.coo_row.p, m.coo_col.p, m.coo_val.p, nnz, row_cur.p, col_cur.p,
                                                                                    m.csr_col.p, m.csr_val.p, m.csc_row.p, m.csc_val.p, cb,
                                                                                    m.csr_code.p, m.csc_code.p);
  const uint32_t cap = 4096;
  m.long_cols.alloc(cap, st);
  n_long.zero();
  ++g_kernel_launches, k_find_long_cols<<<(unsigned)((num_cols + 255) / 256), 256, 0, st>>>(m.csc_ptr.p, num_cols, m.long_cols.p, n_long.p, cap,
                                                                                            (uint32_t)kLongCol);
  uint32_t h_long = 0;
  VPIN_CUDA(cudaMemcpyAsync(&h_long, n_long.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ctx->sync();
  VPIN_REQUIRE(h_long <= cap, VPIN_ERR_BAD_ARGUMENT, "too many dense columns");
  m.n_long = h_long;
  // rows with more than kLongRow entries get a warp each in the SpMV (the capacity grows with the instance: vPIN has one
  // 128-term row per multiplication)
  const uint32_t row_cap = (uint32_t)std::max<size_t>(4096, num_rows / 64);
  m.long_rows.alloc(row_cap, st);
  n_long.zero();
  ++g_kernel_launches, k_find_long_cols<<<(unsigned)((num_rows + 255) / 256), 256, 0, st>>>(m.csr_ptr.p, num_rows, m.long_rows.p, n_long.p, row_cap,
                                                                                            (uint32_t)kLongRow);
  VPIN_CUDA(cudaMemcpyAsync(&h_long, n_long.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ctx->sync();
  // more long rows than the list holds: none of them is listed and the thread-per-row kernel does them all (correct, slower)
  m.n_long_rows = h_long <= row_cap ? h_long : 0;
  return 0;
}
}  // namespace

// Spartan/src/lib.rs:138-244 (Instance::new): padding rules and error behaviour; the zlib digest (:241) is unused on
// the SNARK path and is not computed.
std::unique_ptr<Instance> instance_create(Ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs,
                                          const vpin_coo_entry *A, uint64_t nA, const vpin_coo_entry *B, uint64_t nB,
                                          const vpin_coo_entry *C, uint64_t nC, bool entries_on_device) {
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
    size_t n_pad = (num_cons == 0 || num_cons == 1) && cnt[k] < num_cons_padded ? num_cons_padded - cnt[k] : 0;
    uint32_t err = build_matrix(ctx, inst->M[k], src[k], entries_on_device, cnt[k], n_pad, num_cons, num_vars, num_inputs, num_cons_padded, num_vars_padded);
    VPIN_REQUIRE(err != 2, VPIN_ERR_INVALID_INDEX, "InvalidIndex");
    VPIN_REQUIRE(err != 1, VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
  }
  return inst;
}

}  // namespace vpin
