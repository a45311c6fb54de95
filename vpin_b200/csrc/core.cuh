// Context, HBM buffers, generator streams + fixed-base tables, Hyrax commitment driver, R1CS instance storage.
#pragma once
#include "launch_count.hpp"
#include <cuda_runtime.h>

#include <atomic>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vpin_b200.h"
#include "ed.cuh"
#include "host_fast.hpp"
#include "kernels_msm.cuh"
#include "kernels_poly.cuh"
#include "merlin.hpp"

namespace vpin {

struct Error : std::runtime_error {
  vpin_status code;
  Error(vpin_status c, const std::string &m) : std::runtime_error(m), code(c) {}
};
#define VPIN_CUDA(x)                                                                                       \
  do {                                                                                                     \
    cudaError_t e_ = (x);                                                                                  \
    if (e_ != cudaSuccess)                                                                                 \
      throw ::vpin::Error(e_ == cudaErrorMemoryAllocation ? VPIN_ERR_OOM : VPIN_ERR_CUDA,                  \
                          std::string(#x) + ": " + cudaGetErrorString(e_));                                \
  } while (0)
#define VPIN_REQUIRE(cond, code, msg)                    \
  do {                                                   \
    if (!(cond)) throw ::vpin::Error((code), (msg));     \
  } while (0)


// Per-stream caching allocator. Every context drives ONE stream from one host thread, so a block released by a DevVec may be
// handed to the next DevVec on the same stream without any synchronisation (stream order protects it). Blocks come from
// cudaMalloc, are kept in exact-size free lists and only go back to the driver when the context dies: a steady-state
// proof makes no driver allocation at all. (cudaMallocAsync was measured to cost 10-450 ms per proof once two contexts
// share the default pool or the pool fragments — its reuse across streams either stalls or maps fresh memory.)
struct BlockCache;
BlockCache *block_cache_of(cudaStream_t st);            // nullptr when the stream has no cache (context gone)
void *block_cache_alloc(BlockCache *c, size_t bytes);   // throws Error(VPIN_ERR_OOM)
void block_cache_free(BlockCache *c, void *p);
void block_cache_register(cudaStream_t st);
void block_cache_unregister(cudaStream_t st);           // frees every cached block; outstanding blocks are freed on release
// `hook` is called (without the cache lock) when the driver is out of memory even after the idle blocks went back to it;
// it releases whatever the owner can spare (the context's idle SPARK workspace) and returns true if that was anything.
void block_cache_set_pressure_hook(cudaStream_t st, std::function<bool()> hook);

// Device-synchronisation gate. cudaFree / cudaFreeHost wait for the whole device to go idle and keep other threads' launches
// out while they do. A pre-launched round kernel (kernels_poly.cuh ChalSlot) idles only after its HOST thread has posted a
// challenge, and that thread may be sitting in exactly such a blocked launch: measured as a 20 s stall (until the mailbox timed
// out) when one context was destroyed while another context's proof was in a pre-launch window. Every freeing call of the
// library therefore takes the gate exclusively and a prover holds it shared while one of its kernels may be waiting for a
// post (prover.cu ChalGuard). Frees of other libraries in the process are not covered: the mailbox time-out and the retry of
// vpin_prove without pre-launch (capi.cu) are the backstop for those.
std::shared_mutex &device_sync_gate();
void gated_cuda_free(void *p);
void gated_cuda_free_host(void *p);

// HBM array owned through the stream's block cache
template <class T>
struct DevVec {
  T *p = nullptr;
  size_t n = 0;
  cudaStream_t st = nullptr;
  DevVec() {}
  DevVec(size_t n_, cudaStream_t st_) { alloc(n_, st_); }
  DevVec(const DevVec &) = delete;
  DevVec &operator=(const DevVec &) = delete;
  DevVec(DevVec &&o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
  DevVec &operator=(DevVec &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; st = o.st; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevVec() { release(); }
  void alloc(size_t n_, cudaStream_t st_) {
    release();
    n = n_; st = st_;
    if (n) p = static_cast<T *>(block_cache_alloc(block_cache_of(st), n * sizeof(T)));
  }
  void release() {
    if (p) block_cache_free(block_cache_of(st), p);
    p = nullptr; n = 0;
  }
  void upload(const T *h, size_t cnt) { VPIN_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, st)); }
  void download(T *h, size_t cnt) const { VPIN_CUDA(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, st)); }
  void zero() { if (n) VPIN_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
};

// host fixed-base scalar multiplication for the handful of generators used by the sigma protocols (host_fast.hpp)
typedef hf::FixedBase HostBase;

// generators of one SHAKE256 stream (Spartan/src/commitments.rs:20-38) with their fixed-base table in HBM
struct LabelGens {
  std::string label;
  size_t n = 0;                 // points derived: stream[0..n)
  std::vector<ge_t> h_pts;      // host copies (extended)
  DevVec<ge_t> d_pts;
  MsmGeom geom = msm_geom(kMsmMinW);  // window width of this stream's table (chosen when the table is built)
  DevVec<niels_t> d_table;      // msm_table_entries(n, geom) entries
  std::map<size_t, std::unique_ptr<HostBase>> host_bases;  // built on first use, kept with the stream
  std::mutex mu;                // a stream's generators are shared by every context of the process on the same device
  MsmTable table() const { return MsmTable{d_table.p, n, geom}; }
  const HostBase *host_base(size_t index) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = host_bases.find(index);
    if (it != host_bases.end()) return it->second.get();
    auto b = std::make_unique<HostBase>();
    b->build(hf::ge_from_dev(h_pts.at(index)));
    return (host_bases[index] = std::move(b)).get();
  }
};

struct vpin_ctx_impl;
typedef vpin_ctx_impl Ctx;

// Per-kernel-class device timing (CUDA events on the context stream) for bench.py's roofline lines.
// Off by default; when on, scopes whose work is below `min_units` are not timed (keeps the event overhead out of the
// latency-bound tail rounds).
enum ProfClass {
  PROF_MSM_RECODE, PROF_MSM_ACCUMULATE, PROF_MSM_FINISH, PROF_SC_CUBIC, PROF_SC_QUAD, PROF_SC_BATCHED, PROF_BIND, PROF_SPMV,
  PROF_SPMV_T, PROF_EQ, PROF_TREE, PROF_HASH, PROF_GATHER, PROF_BOUND, PROF_DOT, PROF_BULLET, PROF_FINAL, PROF_COUNT
};
const char *prof_class_name(int cls);
struct Prof {
  bool on = false;
  double min_units = 32768;
  struct Acc { double ms = 0; uint64_t launches = 0; double units = 0; double bytes = 0; } acc[PROF_COUNT];
  struct Pending { int cls; cudaEvent_t e0, e1; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
};
struct ProfScope {
  Ctx *c;
  int cls;
  cudaEvent_t e0 = nullptr;
  ProfScope(Ctx *ctx, int cls, double units, double bytes, int launches = 1);
  ~ProfScope();
};
void prof_drain(Ctx *ctx);

struct vpin_ctx_impl {
  int device = 0;
  cudaStream_t st = nullptr;
  std::string err;
  std::atomic<uint64_t> kernel_launches{0};  // launches made by C-ABI calls on this context (launch_count.hpp)
  std::map<std::string, std::shared_ptr<LabelGens>> label_gens;
  DevVec<fl_t> d_partials;   // reduction scratch
  DevVec<fl_t> d_small;      // small device results / parameters
  fl_t *h_small = nullptr;   // pinned mirror
  // fused sumcheck rounds (kernels_round.cu): two host-mapped result slots used alternately, the counters of the
  // last-block reductions, and the sequence number of the latest launch
  RoundSlot *h_slots = nullptr, *d_slots = nullptr;
  fl_t *h_tail = nullptr, *d_tail = nullptr;  // host-mapped buffer (kTailElems) for the table heads of a layer's host-finished tail
  DevVec<unsigned> d_round_counters;
  uint32_t round_seq = 0;
  // challenge mailbox of the pre-launched round kernels (kernels_poly.cuh ChalSlot): a ring of host-mapped slots the host posts
  // challenges into, and the device-side latch that relays a posted challenge to the other blocks of a grid
  ChalSlot *h_chal = nullptr, *d_chal = nullptr;
  DevVec<ChalLatch> d_chal_latch;
  // background context (vpin_ctx_create_ex(.., -1, ..)): lowest stream priority and at most one resident MSM block per SM - for
  // work nothing waits for (the commitment half of SNARK::encode while the proof runs on another context of the device)
  bool background = false;
  bool no_prelaunch = false;  // set after a mailbox time-out: the context keeps proving with challenges as kernel parameters
  // per-proof workspace for the SPARK tables (derefs, product trees, dot-product clones): one slab that only ever grows, so
  // a steady-state proof allocates nothing large (multi-GB cudaMallocAsync calls were measured at 10-150 ms when the pool
  // has to grow or is fragmented)
  DevVec<fl_t> workspace;
  bool workspace_busy = false;  // a proof is using the slab (set by the prover): the out-of-memory hook must leave it alone
  fl_t *workspace_reserve(size_t elems) {
    if (workspace.n < elems) {
      workspace.release();
      sync();
      workspace.alloc(elems, st);
    }
    return workspace.p;
  }
  std::vector<std::pair<const char *, double>> phases;
  Prof prof;
  // multi-GPU (one process per GPU): NCCL communicator over NVLink/NVSwitch, created by vpin_ctx_init_distributed
  int rank = 0, world = 1;
  void *nccl_comm = nullptr;
  std::shared_ptr<void> host_pool;  // the prover's helper threads (prover.cu HostPool), created with the first proof
  // A second stream for work of one proof that does not depend on the transcript's next challenges (the row half of the derefs
  // commitment runs under the second sumcheck): SideScope (below) forks it off the main stream and swaps it in as `st`.
  cudaStream_t st_main = nullptr, st_side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool on_side = false;
  int shard_sumcheck = -1;  // sharded sumcheck rounds of one proof: -1 = VPIN_SHARD_SUMCHECK decides, 0 / 1 = vpin_ctx_set_shard_sumcheck
  DevVec<unsigned long long> d_counters;  // [0] = non-zero MSM digits recoded (= mixed additions executed)
  // sharded sumcheck rounds (one proof on several GPUs, VPIN_SHARD_SUMCHECK=1): device-side result slots of the round kernels
  // and the all-gather buffer (kRoundSlotVals elements)
  DevVec<RoundSlot> d_dev_slots;
  DevVec<fl_t> d_gather;
  // drains the stream. On a distributed context the wait is a poll with a deadline (VPIN_DIST_TIMEOUT_S, default 120 s): a peer
  // that failed never joins the collective this stream may be waiting in; the communicator is then aborted (dist_abort) and the
  // call fails instead of hanging.
  void sync() {
    if (world > 1 && nccl_comm) sync_distributed();
    else VPIN_CUDA(cudaStreamSynchronize(st));
  }
  void sync_distributed();
  // a marker on the stream that the host can wait for without draining what is queued behind it
  cudaEvent_t ev_marker = nullptr;
  void mark() {
    if (!ev_marker) VPIN_CUDA(cudaEventCreateWithFlags(&ev_marker, cudaEventDisableTiming));
    VPIN_CUDA(cudaEventRecord(ev_marker, st));
  }
  void wait_mark() { VPIN_CUDA(cudaEventSynchronize(ev_marker)); }
};

// While alive, everything the context enqueues (kernels, DevVec allocations, profiling events) goes to the side stream, which
// starts after what the main stream holds now; join() makes the main stream wait for it. Not for distributed contexts (the
// communicator's collectives must stay on one stream).
int side_msm_blocks_per_sm();
struct SideScope {
  Ctx *c;
  explicit SideScope(Ctx *ctx);
  ~SideScope();
  static void join(Ctx *ctx);  // main stream waits for everything the last SideScope enqueued
};
// rows [*r0, *r1) of an L-row Hyrax grid that rank `rank` of `world` commits to; the whole range when the grid is too
// small to shard (fewer than kMinShardRows rows per rank) or not divisible. Returns true when the grid is sharded.
static const size_t kMinShardRows = 32;
bool shard_rows(size_t rows, int rank, int world, size_t *r0, size_t *r1);
// in-place all-gather of `bytes_per_rank` bytes per rank inside buf (rank r's slice at offset r * bytes_per_rank)
void dist_allgather_inplace(Ctx *ctx, void *buf, size_t bytes_per_rank);
void dist_broadcast_many(Ctx *ctx, void *const *bufs, const size_t *bytes, const int *roots, int count);
void dist_get_unique_id(uint8_t out[128]);
void dist_init(Ctx *ctx, int rank, int world, const uint8_t id[128]);
void dist_destroy(Ctx *ctx);
void dist_abort(Ctx *ctx);  // ncclCommAbort: releases a collective a failed peer will never join

// table_budget: bytes the fixed-base table may take (0 = 30 % of the free HBM). Generators and tables are public parameters:
// one copy per (device, label) serves every context of the process (a second context proving the same shape neither
// rebuilds nor duplicates 30 GB of tables); *built tells whether this call had to build one.
std::shared_ptr<LabelGens> get_label_gens(Ctx *ctx, const std::string &label, size_t n, size_t table_budget = 0, bool *built = nullptr);
std::shared_ptr<LabelGens> find_label_gens(Ctx *ctx, const std::string &label, size_t n);  // nullptr when none is cached yet

// Hyrax rows: out[i] = sum_j Z[i*ld + j] * G_j (+ blind_i * G_{blind_base}). d_points (rows ge_t) and d_comp (rows*32 B)
// are optional outputs.
void hyrax_rows(Ctx *ctx, const LabelGens &g, const fl_t *dZ, size_t rows, size_t cols, size_t ld, const fl_t *d_blinds,
                size_t blind_base, ge_t *d_points, uint8_t *d_comp);

// one sparse matrix of the instance in the three layouts the kernels use
struct MatrixDev {
  size_t nnz = 0;
  DevVec<uint32_t> coo_row, coo_col;    // COO order == SPARK "ops" order (Spartan/src/sparse_mlpoly.rs:368-380)
  DevVec<fl_t> coo_val;
  DevVec<uint32_t> csr_ptr, csr_col;
  DevVec<fl_t> csr_val;
  DevVec<uint32_t> csc_ptr, csc_row, long_cols;
  DevVec<fl_t> csc_val;
  size_t n_long = 0;
  // value dictionary (vPIN's R1CS coefficients are +-1, +-2, 3 except for the 2^i of the bit decompositions):
  // one byte per CSR / CSC entry (kernels_poly.cuh SpmvCode); the SpMV kernels only read the 32-byte value when the code says so
  DevVec<uint8_t> csr_code, csc_code;
  DevVec<uint32_t> long_rows;  // rows with more than kLongRow entries (one 128-term bit decomposition per multiplication)
  size_t n_long_rows = 0;
};
struct vpin_instance_impl {
  size_t num_cons = 0, num_vars = 0, num_inputs = 0;  // padded cons / vars (Spartan/src/lib.rs:146-176)
  MatrixDev M[3];
};
typedef vpin_instance_impl Instance;

std::unique_ptr<Instance> instance_create(Ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs,
                                          const vpin_coo_entry *A, uint64_t nA, const vpin_coo_entry *B, uint64_t nB,
                                          const vpin_coo_entry *C, uint64_t nC, bool entries_on_device = false);

static inline size_t log2_ceil(size_t x) { size_t l = 0; while (((size_t)1 << l) < x) l++; return l; }
static inline size_t next_pow2(size_t x) { return (size_t)1 << log2_ceil(x); }

}  // namespace vpin
