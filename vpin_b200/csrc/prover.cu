// B200 prover for vPIN's my_lib_prove / SNARK::encode. "SP/" = Spartan/src/, "VP/" = vPIN_proof_generation/src/.
// Host side: Merlin transcript, random tape, O(1)-size sigma protocols, bincode. Device side: everything whose cost
// grows with the instance (tables in HBM, see kernels_poly.cu / kernels_msm.cu).
#include "prover.cuh"
#include <shared_mutex>

#include <time.h>

#include <array>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace vpin {

typedef std::array<uint8_t, 32> Comp;

static double now_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
typedef hf::ge hge_t;  // host-side points (5 x 51-bit limbs, host_fast.hpp)
static Comp compress_host(const hge_t &g) {
  Comp c;
  hf::ge_compress(g, c.data());
  return c;
}
// Math::log_2 (SP/math.rs:27-35): exact for powers of two, ceil otherwise
static size_t math_log2(size_t x) {
  VPIN_REQUIRE(x != 0, VPIN_ERR_BAD_ARGUMENT, "log_2(0)");
  return log2_ceil(x);
}

// ------------------------------------------------------------------------------------------------ bincode 1.3.3
struct Bin {
  std::vector<uint8_t> b;
  void u64(uint64_t x) { b.insert(b.end(), (uint8_t *)&x, (uint8_t *)&x + 8); }
  void fl(const fl_t &x) { b.insert(b.end(), (const uint8_t *)x.v, (const uint8_t *)x.v + 32); }  // raw Montgomery limbs
  void comp(const Comp &c) { b.insert(b.end(), c.begin(), c.end()); }
  void fls(const std::vector<fl_t> &v) { u64(v.size()); for (auto &x : v) fl(x); }
  void comps(const std::vector<Comp> &v) { u64(v.size()); for (auto &x : v) comp(x); }
};

// proof pieces in the reference's field order
struct DotProductProofS { Comp delta, beta; std::vector<fl_t> z; fl_t z_delta, z_beta; };
static void put(Bin &o, const DotProductProofS &p) { o.comp(p.delta); o.comp(p.beta); o.fls(p.z); o.fl(p.z_delta); o.fl(p.z_beta); }
struct ZkSumcheckS { std::vector<Comp> comm_polys, comm_evals; std::vector<DotProductProofS> proofs; };
static void put(Bin &o, const ZkSumcheckS &p) {
  o.comps(p.comm_polys); o.comps(p.comm_evals);
  o.u64(p.proofs.size());
  for (auto &x : p.proofs) put(o, x);
}
struct KnowledgeS { Comp alpha; fl_t z1, z2; };
static void put(Bin &o, const KnowledgeS &p) { o.comp(p.alpha); o.fl(p.z1); o.fl(p.z2); }
struct EqualityS { Comp alpha; fl_t z; };
static void put(Bin &o, const EqualityS &p) { o.comp(p.alpha); o.fl(p.z); }
struct ProductS { Comp alpha, beta, delta; fl_t z[5]; };
static void put(Bin &o, const ProductS &p) { o.comp(p.alpha); o.comp(p.beta); o.comp(p.delta); for (int i = 0; i < 5; i++) o.fl(p.z[i]); }
struct DotLogS { std::vector<Comp> L_vec, R_vec; Comp delta, beta; fl_t z1, z2; };  // PolyEvalProof { DotProductProofLog }
static void put(Bin &o, const DotLogS &p) { o.comps(p.L_vec); o.comps(p.R_vec); o.comp(p.delta); o.comp(p.beta); o.fl(p.z1); o.fl(p.z2); }
struct LayerS { std::vector<std::vector<fl_t>> polys; std::vector<fl_t> left, right; };  // LayerProofBatched
struct BatchedS { std::vector<LayerS> layers; std::vector<fl_t> dotp[3]; };             // ProductCircuitEvalProofBatched
static void put(Bin &o, const BatchedS &p) {
  o.u64(p.layers.size());
  for (auto &l : p.layers) {
    o.u64(l.polys.size());
    for (auto &c : l.polys) o.fls(c);
    o.fls(l.left);
    o.fls(l.right);
  }
  o.fls(p.dotp[0]); o.fls(p.dotp[1]); o.fls(p.dotp[2]);
}

// ------------------------------------------------------------------------------------------------ gens
static void make_pc(Ctx *ctx, LabelGens &lg, size_t ell, PcGens *pc) {
  pc->ell = ell;
  pc->L = (size_t)1 << (ell / 2);
  pc->R = (size_t)1 << (ell - ell / 2);
  pc->g1_index = pc->R;
  pc->h_index = pc->R + 1;
  VPIN_REQUIRE(pc->h_index < lg.n, VPIN_ERR_SIZE_MISMATCH, "generator stream too short");
  pc->g1 = lg.host_base(pc->g1_index);
  pc->h = lg.host_base(pc->h_index);
}
// SP/lib.rs:305-326, SP/r1csproof.rs:84-89, SP/r1csinstance.rs:35-48, SP/sparse_mlpoly.rs:302-327
std::unique_ptr<SnarkGens> snark_gens_create(Ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs, uint64_t num_nz_entries) {
  size_t num_vars_padded = next_pow2(std::max<size_t>(num_vars, num_inputs + 1));
  VPIN_REQUIRE(num_inputs < num_vars_padded, VPIN_ERR_INVALID_NUM_INPUTS, "num_inputs must be < num_vars");
  VPIN_REQUIRE(num_nz_entries > 0, VPIN_ERR_BAD_ARGUMENT, "num_nz_entries must be > 0");
  auto g = std::make_unique<SnarkGens>();
  size_t ell_sat = math_log2(num_vars_padded);
  size_t R_sat = (size_t)1 << (ell_sat - ell_sat / 2);
  size_t nvx = math_log2(num_cons), nvy = math_log2(2 * num_vars_padded);
  size_t k = math_log2(next_pow2(num_nz_entries));
  size_t ell_ops = k + math_log2(next_pow2(3 * 5)), ell_mem = std::max(nvx, nvy) + 1, ell_derefs = k + math_log2(next_pow2(3 * 2));
  size_t ell_max = std::max(ell_ops, std::max(ell_mem, ell_derefs));
  size_t R_max = (size_t)1 << (ell_max - ell_max / 2);
  // HBM plan for this shape: what one proof keeps alive next to the two generator tables — the SPARK workspace slab
  // (41 N + 8 M elements, snark_prove), the decommitment (comb_ops 16 N, comb_mem 2 M, address / timestamp words), the
  // instance (COO + CSR + CSC), the witness-sized tables of the satisfiability proof, MSM / sort temporaries — plus a
  // margin. The tables share what is left: 80 % at most for the evaluation generators (derefs 8 N and comb_ops 16 N
  // scalars go through them), the rest for the satisfiability generators. LeNet layer 5 (N = 2^25) keeps ~115 GB alive:
  // without the plan the first table took 30 % of an empty GPU and the second proof ran out of memory.
  // (the plan - one cudaMemGetInfo, tens of milliseconds on a full device - is only drawn up when a table has to be built)
  const size_t n_eval = R_max + 2, n_sat = std::max<size_t>(R_sat + 2, 5);
  g->eval_label = find_label_gens(ctx, "gens_r1cs_eval", n_eval);
  g->sat_label = find_label_gens(ctx, "gens_r1cs_sat", n_sat);
  if (!g->eval_label || !g->sat_label) {
    size_t N = (size_t)1 << k, M = (size_t)1 << std::max(nvx, nvy), c = (size_t)1 << nvx, v = num_vars_padded;
    size_t working = (41 * N + 8 * M) * 32 + (16 * N + 2 * M) * 32 + 64 * N + 3 * N * 112 + (8 * v + 5 * c) * 32 + ((size_t)4 << 30);
    size_t free_b = 0, total_b = 0;
    VPIN_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t margin = (size_t)6 << 30;
    size_t avail = free_b > working + margin ? free_b - working - margin : 0;
    size_t table_budget_eval = std::max<size_t>(avail / 5 * 4, 1);
    bool eval_built = false;
    if (!g->eval_label) g->eval_label = get_label_gens(ctx, "gens_r1cs_eval", n_eval, table_budget_eval, &eval_built);
    size_t eval_bytes = eval_built ? msm_table_entries(g->eval_label->n, g->eval_label->geom) * sizeof(niels_t) : 0;
    size_t table_budget_sat = std::max<size_t>(avail > eval_bytes ? avail - eval_bytes : 0, 1);
    if (!g->sat_label) g->sat_label = get_label_gens(ctx, "gens_r1cs_sat", n_sat, table_budget_sat);
  }
  make_pc(ctx, *g->sat_label, ell_sat, &g->sat_pc);
  for (int i = 0; i < 5; i++) g->sat_g[i] = g->sat_label->host_base(i);
  make_pc(ctx, *g->eval_label, ell_ops, &g->ops_pc);
  make_pc(ctx, *g->eval_label, ell_mem, &g->mem_pc);
  make_pc(ctx, *g->eval_label, ell_derefs, &g->derefs_pc);
  return g;
}

// ------------------------------------------------------------------------------------------------ encode
// SP/lib.rs:347-358 -> SP/r1csinstance.rs:309-321 -> SP/sparse_mlpoly.rs:500-520, :382-438, AddrTimestamps::new :232-265
std::unique_ptr<Decomm> snark_encode(Ctx *ctx, const Instance &inst, const SnarkGens &gens, std::vector<uint8_t> *comm_bytes) {
  auto d = snark_encode_tables(ctx, inst, gens);
  *comm_bytes = snark_encode_commit(ctx, *d, gens);
  return d;
}
std::unique_ptr<Decomm> snark_encode_tables(Ctx *ctx, const Instance &inst, const SnarkGens &gens) {
  cudaStream_t st = ctx->st;
  auto d = std::make_unique<Decomm>();
  d->num_cons = inst.num_cons; d->num_vars = inst.num_vars; d->num_inputs = inst.num_inputs;
  size_t N = 1;
  for (int k = 0; k < 3; k++) N = std::max(N, next_pow2(inst.M[k].nnz));
  size_t nvx = math_log2(inst.num_cons), nvy = math_log2(2 * inst.num_vars);
  size_t M = (size_t)1 << std::max(nvx, nvy);
  d->N = N;
  d->M = M;
  VPIN_REQUIRE(math_log2(16 * N) == gens.ops_pc.ell && math_log2(2 * M) == gens.mem_pc.ell, VPIN_ERR_SIZE_MISMATCH,
               "gens do not match the instance (SP/commitments.rs:95 assert)");
  // read / audit timestamps (:237-257; audit counters shared by the three matrices) from the COO index arrays already in
  // HBM: stable sort by address instead of the reference's sequential replay (kernels_sort.cu)
  d->comb_ops.alloc(16 * N, st);
  d->comb_ops.zero();
  d->row_audit_ts.alloc(M, st); d->col_audit_ts.alloc(M, st);
  {
    DevVec<uint32_t> scratch(spark_timestamps_scratch_words(N, M), st);
    size_t nnz[3] = {inst.M[0].nnz, inst.M[1].nnz, inst.M[2].nnz};
    for (int pass = 0; pass < 2; pass++) {
      DevVec<uint32_t> &da = pass == 0 ? d->row_addr_all : d->col_addr_all;
      DevVec<uint32_t> &dt = pass == 0 ? d->row_read_ts_all : d->col_read_ts_all;
      da.alloc(3 * N, st); dt.alloc(3 * N, st);
      const uint32_t *addr[3];
      for (int k = 0; k < 3; k++) addr[k] = pass == 0 ? inst.M[k].coo_row.p : inst.M[k].coo_col.p;
      launch_spark_timestamps(addr, nnz, N, M, da.p, dt.p, pass == 0 ? d->row_audit_ts.p : d->col_audit_ts.p, scratch.p, st);
      for (int k = 0; k < 3; k++) {
        (pass == 0 ? d->row_addr : d->col_addr)[k].p = da.p + (size_t)k * N;
        (pass == 0 ? d->row_read_ts : d->col_read_ts)[k].p = dt.p + (size_t)k * N;
      }
      launch_u32_to_fl(da.p, 3 * N, d->comb_ops.p + (size_t)(pass * 6) * N, st);
      launch_u32_to_fl(dt.p, 3 * N, d->comb_ops.p + (size_t)(pass * 6 + 3) * N, st);
    }
  }
  for (int k = 0; k < 3; k++)
    if (inst.M[k].nnz)
      VPIN_CUDA(cudaMemcpyAsync(d->comb_ops.p + (12 + k) * N, inst.M[k].coo_val.p, inst.M[k].nnz * sizeof(fl_t), cudaMemcpyDeviceToDevice, st));
  d->comb_mem.alloc(2 * M, st);
  launch_u32_to_fl(d->row_audit_ts.p, M, d->comb_mem.p, st);
  launch_u32_to_fl(d->col_audit_ts.p, M, d->comb_mem.p + M, st);
  ctx->sync();  // the tables may be read from any stream once this returns
  return d;
}
std::vector<uint8_t> snark_encode_commit(Ctx *ctx, const Decomm &dec, const SnarkGens &gens) {
  cudaStream_t st = ctx->st;
  const Decomm *d = &dec;
  const size_t N = dec.N, M = dec.M;
  VPIN_REQUIRE(math_log2(16 * N) == gens.ops_pc.ell && math_log2(2 * M) == gens.mem_pc.ell, VPIN_ERR_SIZE_MISMATCH,
               "gens do not match the decommitment");
  // two Hyrax commitments without blinds (:507-508)
  const PcGens &po = gens.ops_pc, &pm = gens.mem_pc;
  DevVec<uint8_t> c_ops(32 * po.L, st), c_mem(32 * pm.L, st);
  hyrax_rows(ctx, *gens.eval_label, d->comb_ops.p, po.L, po.R, po.R, nullptr, 0, nullptr, c_ops.p);
  hyrax_rows(ctx, *gens.eval_label, d->comb_mem.p, pm.L, pm.R, pm.R, nullptr, 0, nullptr, c_mem.p);
  std::vector<Comp> h_ops(po.L), h_mem(pm.L);
  c_ops.download((uint8_t *)h_ops.data(), 32 * po.L);
  c_mem.download((uint8_t *)h_mem.data(), 32 * pm.L);
  ctx->sync();
  // bincode(ComputationCommitment { comm: R1CSCommitment { num_cons, num_vars, num_inputs, comm: SparseMatPolyCommitment {
  //   batch_size, num_ops, num_mem_cells, comm_comb_ops, comm_comb_mem } } })   SP/r1csinstance.rs:53-58, sparse_mlpoly.rs:332-338
  Bin o;
  o.u64(dec.num_cons); o.u64(dec.num_vars); o.u64(dec.num_inputs);
  o.u64(3); o.u64(N); o.u64(M);
  o.comps(h_ops);
  o.comps(h_mem);
  return std::move(o.b);
}

// ------------------------------------------------------------------------------------------------ prover
namespace {

// ONE proof on several GPUs beyond the commitment rows: the product circuits of the SPARK memory check (12 over the N ops, 4
// over the M memory cells, + 6 dot-product instances) are DEALT to the ranks - a rank builds the hash vectors and product trees
// of its own circuits only, and runs only their instances in the batched sumchecks of the large layers (those of at least
// kShardMinLayer thread items; one small NCCL all-gather per round, see batched_prove); the short upper layers of every tree are
// broadcast once by their owners and proved replicated. Mid-size circuit sets (fewer than kDealBuildMin leaves) build every tree
// everywhere and deal only the rounds of their largest layers (>= kRoundDealMin items): measured on 8 B200s, CNN A loses 10 ms
// when more is dealt (a per-round exchange costs ~45 us, the broadcast of 16 tree tails 0.5 ms). Inside a sharded layer the
// rounds go back to replicated once they are short (kUnshardQ items; the owners broadcast their bound tables). On by default for
// a distributed context (VPIN_SHARD_SUMCHECK=0 or vpin_ctx_set_shard_sumcheck(ctx, 0) turn it off); the proof bytes do not change.
static const size_t kShardMinLayer = (size_t)1 << 17;  // smallest sharded layer of a circuit set whose BUILD is dealt (thread items)
static const size_t kRoundDealMin = (size_t)1 << 19;   // smallest sharded layer when only the rounds are dealt (trees replicated)
static const size_t kDealBuildMin = (size_t)1 << 22;   // leaves from which hash vectors and trees are built by their owners only
static const size_t kUnshardQ = (size_t)1 << 11;       // a sharded layer goes back to replicated rounds at this many thread items
// The tail of a batched layer on the host: once a round has at most this many thread items the heads of all its tables (4 q
// elements each) travel to the host in ONE copy and the host binds, evaluates and delivers the final claims itself - ~60 field
// multiplications per instance instead of 3 - 4 more launch / PCIe round trips of ~20 us each (VPIN_HOST_TAIL_Q overrides, 0 = off)
static const size_t kHostTailQDefault = 4;
static bool shard_sumcheck_enabled(const Ctx *ctx) {
  static const bool env_on = [] { const char *e = getenv("VPIN_SHARD_SUMCHECK"); return !e || atoi(e) != 0; }();
  const bool on = ctx->shard_sumcheck < 0 ? env_on : ctx->shard_sumcheck != 0;
  return on && ctx->world > 1 && ctx->nccl_comm != nullptr;
}

// A few helper threads for the O(1) elliptic-curve work next to the transcript (the per-round commitments of the ZK
// sumchecks, the L / R points of a bullet-reduction round): ~10 fixed-base multiplications of 5-10 us each per round that
// are independent of one another. Helpers spin while a scope is active (hand-off ~0.3 us) and sleep otherwise.
class HostPool {
 public:
  explicit HostPool(int n) : slots_(n) {
    for (int i = 0; i < n; i++) th_.emplace_back([this, i] { loop(i); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_.store(true);
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  int size() const { return (int)slots_.size(); }
  void set_active(bool on) {
    {
      std::lock_guard<std::mutex> g(mu_);
      active_.store(on);
    }
    if (on) cv_.notify_all();
  }
  template <class F>
  void run(int i, F &&f) {
    if (slots_.empty()) { f(); return; }  // no helpers (VPIN_HOST_HELPERS=0, or few cores per rank): do it here
    if (!active_.load(std::memory_order_acquire)) { f(); return; }  // no Scope open: the helpers sleep and would never pick it up
    Slot &s = slots_[i % slots_.size()];
    if (s.state.load(std::memory_order_acquire) != 0) { f(); return; }  // slot busy (fewer helpers than tasks)
    s.fn = std::forward<F>(f);
    s.state.store(1, std::memory_order_release);
  }
  void wait(int i) {
    if (slots_.empty()) return;
    Slot &s = slots_[i % slots_.size()];
    if (s.state.load(std::memory_order_acquire) == 0) return;  // ran inline
    while (s.state.load(std::memory_order_acquire) != 2) __builtin_ia32_pause();
    s.state.store(0, std::memory_order_relaxed);
  }
  // helper threads per prover: VPIN_HOST_HELPERS (default 2: measured 1.3 ms faster per CNN-A proof than none; a third
  // one adds nothing measurable and makes a 16-core host more sensitive to scheduling noise)
  static int default_size() {
    int n = 2;
    if (const char *e = getenv("VPIN_HOST_HELPERS")) n = atoi(e);
    unsigned hc = std::thread::hardware_concurrency();
    if (hc && hc < 8) n = 0;
    return n < 0 ? 0 : (n > 3 ? 3 : n);
  }
  struct Scope {  // helpers spin for the lifetime of the scope
    HostPool &p;
    explicit Scope(HostPool &p_) : p(p_) { p.set_active(true); }
    ~Scope() { p.set_active(false); }
  };
  // Declared AFTER the locals a posted task captures by reference: on any exit from the block - normal or by exception - every
  // task has finished before those locals die (wait() is a no-op for a slot that was already waited for).
  struct Joiner {
    HostPool &p;
    explicit Joiner(HostPool &p_) : p(p_) {}
    ~Joiner() { for (int i = 0; i < p.size(); i++) p.wait(i); }
  };

 private:
  struct alignas(64) Slot {
    std::atomic<int> state{0};  // 0 idle, 1 task posted, 2 task done
    std::function<void()> fn;
  };
  void loop(int i) {
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return active_.load() || stop_.load(); });
      }
      if (stop_.load()) return;
      while (active_.load(std::memory_order_relaxed) && !stop_.load(std::memory_order_relaxed)) {
        if (slots_[i].state.load(std::memory_order_acquire) == 1) {
          slots_[i].fn();
          slots_[i].state.store(2, std::memory_order_release);
        } else {
          __builtin_ia32_pause();
        }
      }
      // a task posted just before the scope closed must still run
      if (slots_[i].state.load(std::memory_order_acquire) == 1) {
        slots_[i].fn();
        slots_[i].state.store(2, std::memory_order_release);
      }
    }
  }
  std::vector<Slot> slots_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<bool> active_{false}, stop_{false};
};

static HostPool &host_pool_of(Ctx *ctx) {
  if (!ctx->host_pool) ctx->host_pool = std::shared_ptr<void>(new HostPool(HostPool::default_size()), [](void *p) { delete static_cast<HostPool *>(p); });
  return *static_cast<HostPool *>(ctx->host_pool.get());
}

struct Prover {
  Ctx *ctx;
  cudaStream_t st;
  MerlinTranscript &t;
  ProverTape &tape;
  const SnarkGens &g;
  HostPool &pool;  // one per context (created with the first proof), not one per proof: no thread spawn / join per proof
  size_t ring = 0;
  // accumulating wall-clock timers (reported next to the phases)
  double t_bullet_gpu = 0, t_bullet_host = 0, t_bullet_pre = 0, t_b_wait = 0, t_b_host = 0, t_b_launch = 0, t_b_small_wait = 0;
  size_t n_b_small = 0, n_b_rounds = 0;

  // ---- tiny host <-> device traffic (challenges in, round sums out) ----
  const fl_t *up(const fl_t *vals, size_t n) {  // returns the device address of n freshly uploaded elements
    if (ring + n > 200) ring = 0;  // slots 240.. of d_small hold round results
    memcpy(ctx->h_small + ring, vals, n * sizeof(fl_t));
    fl_t *dst = ctx->d_small.p + ring;
    VPIN_CUDA(cudaMemcpyAsync(dst, ctx->h_small + ring, n * sizeof(fl_t), cudaMemcpyHostToDevice, st));
    ring += n;
    return dst;
  }
  const fl_t *up(const std::vector<fl_t> &v) { return up(v.data(), v.size()); }
  void down(const void *d, size_t bytes, void *out) {
    VPIN_CUDA(cudaMemcpyAsync(ctx->h_small + 256, d, bytes, cudaMemcpyDeviceToHost, st));
    ctx->sync();
    memcpy(out, ctx->h_small + 256, bytes);
  }
  fl_t down1(const fl_t *d) { fl_t x; down(d, sizeof(fl_t), &x); return x; }

  DevVec<fl_t> eq_table(const std::vector<fl_t> &r) {  // the point travels as a kernel parameter: no copy, no sync
    size_t ell = r.size();
    VPIN_REQUIRE(ell <= 32, VPIN_ERR_BAD_ARGUMENT, "eq table: more than 32 variables");
    DevVec<fl_t> out((size_t)1 << ell, st), tmp(eq_tmp_elems(ell), st);
    EqPoint pt;
    for (size_t i = 0; i < ell; i++) pt.r[i] = r[i];
    ProfScope ps(ctx, PROF_EQ, (double)((size_t)1 << ell), 32.0 * (double)((size_t)1 << ell), ell <= 12 ? 1 : 3);
    launch_eq_evals_pt(pt, (int)ell, out.p, tmp.p, st);
    return out;
  }
  // suffix tables of eq(r, .) for the sumchecks that factor the eq polynomial out (launch_eq_suffix): table k at offset 2^k
  DevVec<fl_t> eq_suffix(const std::vector<fl_t> &r) {
    size_t ell = r.size();
    VPIN_REQUIRE(ell <= 32, VPIN_ERR_BAD_ARGUMENT, "eq table: more than 32 variables");
    int kmax = ell ? (int)ell - 1 : 0;
    DevVec<fl_t> out((size_t)2 << kmax, st);
    EqPoint pt;
    for (size_t i = 0; i < ell; i++) pt.r[i] = r[i];
    ProfScope ps(ctx, PROF_EQ, (double)((size_t)1 << kmax), 32.0 * (double)((size_t)1 << kmax), 1 + (kmax > 11 ? (kmax - 11 + 2) / 3 : 0));
    launch_eq_suffix(pt, (int)ell, kmax, out.p, st);
    return out;
  }
  // inverses of every element (Montgomery's trick: one inversion). A zero has no inverse: the split-eq recovery of a round
  // polynomial divides by the coordinates of the verifier's random point, which vanish with probability 2^-252 each.
  static std::vector<fl_t> batch_invert(const std::vector<fl_t> &v) {
    std::vector<fl_t> pre(v.size()), out(v.size());
    fl_t run = fl_one();
    for (size_t i = 0; i < v.size(); i++) {
      VPIN_REQUIRE(!fl_is_zero(v[i]), VPIN_ERR_PROVER, "a coordinate of the random evaluation point is zero");
      pre[i] = run;
      run = run * v[i];
    }
    fl_t inv = fl_invert(run);
    for (size_t i = v.size(); i-- > 0;) {
      out[i] = inv * pre[i];
      inv = inv * v[i];
    }
    return out;
  }
  // One round of a sumcheck whose polynomial has the factor eq(w, .) split off: s(X) = E l(X) t(X), l(X) = (1 - w)(1 - X) + w X,
  // t quadratic. Given t(0), the leading coefficient a = t(inf) and the scaled claim sc = s(0)/E + s(1)/E = l(0) t(0) + l(1) t(1),
  // returns s(0), s(2), s(3) - the evaluations the reference computes table by table (SP/sumcheck.rs:287-357, :619-652) - and
  // keeps what the next round needs. Exact field arithmetic: the same polynomial, hence the same bytes.
  struct SplitEq {
    fl_t t0, t1, a, w;
    static void evals(const fl_t &E, const fl_t &w, const fl_t &winv, const fl_t &sc, const fl_t &t0, const fl_t &a, fl_t out[3], SplitEq *keep) {
      const fl_t one = fl_one();
      fl_t l0 = one - w;
      fl_t t1 = (sc - l0 * t0) * winv;
      fl_t a2 = a + a, t2 = t1 + t1 - t0 + a2, t3 = t2 + t1 - t0 + a2 + a2;  // t(2) = 2 t1 - t0 + 2a, t(3) = 3 t1 - 2 t0 + 6a
      fl_t w2 = w + w, l2 = w2 + w - one, l3 = l2 + w2 - one;                 // l(2) = 3w - 1, l(3) = 5w - 2
      out[0] = E * l0 * t0;
      out[1] = E * l2 * t2;
      out[2] = E * l3 * t3;
      keep->t0 = t0; keep->t1 = t1; keep->a = a; keep->w = w;
    }
    fl_t l_at(const fl_t &r) const { fl_t l0 = fl_one() - w; return l0 + r * (w - l0); }   // eq(w, r)
    fl_t t_at(const fl_t &r) const { return t0 + r * ((t1 - t0 - a) + r * a); }          // the next scaled claim
  };
  // ---- challenge mailbox of the pre-launched rounds (kernels_poly.cuh ChalSlot) ----
  // eight 8-byte atoms (limb, tag); x86 keeps the stores in order and each atom validates itself on the device side
  void chal_write(uint32_t tag, const fl_t &r) {
    volatile uint64_t *w = reinterpret_cast<volatile uint64_t *>(ctx->h_chal + tag % kChalRing);
    for (int k = 0; k < 8; k++) w[k] = (uint64_t)r.v[k] | ((uint64_t)tag << 32);
  }
  // Tags some enqueued kernel still waits for; an unwinding layer releases them with a dummy value. From the first pre-launch of
  // a layer until the guard dies the device-synchronisation gate (core.cuh) is held shared: no cudaFree of this library can
  // stall this thread's launches while a kernel depends on this thread's next post.
  struct ChalGuard {
    Prover *P;
    std::vector<uint32_t> tags;
    std::shared_lock<std::shared_mutex> gate;
    ~ChalGuard() {
      for (uint32_t tg : tags) P->chal_write(tg, fl_zero());
    }
  };
  void chal_post(uint32_t tag, const fl_t &r, ChalGuard *g) {
    if (!test_drop_post()) chal_write(tag, r);
    for (size_t i = 0; i < g->tags.size(); i++)
      if (g->tags[i] == tag) { g->tags.erase(g->tags.begin() + i); break; }
  }
  // how long a pre-launched kernel waits for its challenge before it gives up (VPIN_MAILBOX_TIMEOUT_MS, default 2 s; the proof
  // is then redone without pre-launch, capi.cu)
  static uint32_t mailbox_timeout_ms() {
    static const uint32_t v = [] {
      const char *e = getenv("VPIN_MAILBOX_TIMEOUT_MS");
      long x = e ? atol(e) : 2000;
      return (uint32_t)(x < 1 ? 1 : (x > 600000 ? 600000 : x));
    }();
    return v;
  }
  // test hook: VPIN_TEST_DROP_POST=n loses the n-th post of the process (a host that never answers), once
  static bool test_drop_post() {
    static std::atomic<long> left{[] { const char *e = getenv("VPIN_TEST_DROP_POST"); return e ? atol(e) : 0; }()};
    if (left.load(std::memory_order_relaxed) <= 0) return false;
    return left.fetch_sub(1) == 1;
  }
  // largest round (thread items) whose kernel is enqueued before its challenge exists; VPIN_PRELAUNCH_Q=0 turns pre-launch off
  static size_t prelaunch_q() {
    static const size_t v = [] {
      const char *e = getenv("VPIN_PRELAUNCH_Q");
      // Nsight Compute serialises kernels and holds the launching thread until each one has finished: a kernel that waits for
      // that thread's next post could only time out. Its injection sets NV_COMPUTE_PROFILER_PERFWORKS_DIR in the target process.
      if (!e && getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) return (size_t)0;
      const char *blocking = getenv("CUDA_LAUNCH_BLOCKING");  // every launch then waits for its kernel: same situation
      if (!e && blocking && atoi(blocking) != 0) return (size_t)0;
      long long x = e ? atoll(e) : (1ll << 14);
      return (size_t)(x < 0 ? 0 : x);
    }();
    return v;
  }
  // ---- fused rounds: results arrive in host-mapped slots (kernels_round.cu) ----
  RoundCtl round_ctl(int slot, uint32_t *seq_out) {
    uint32_t seq = ++ctx->round_seq;
    *seq_out = seq;
    return RoundCtl{ctx->d_partials.p, ctx->d_round_counters.p, ctx->d_slots + slot, seq};
  }
  // On a distributed context a result can also stay away because a PEER failed and never joined the collective this rank's
  // stream is waiting in (cudaStreamQuery then says "not ready" forever): after VPIN_DIST_TIMEOUT_S seconds (default 120) the
  // communicator is aborted and the call fails instead of hanging.
  static double dist_timeout_ms() {
    static const double v = [] { const char *e = getenv("VPIN_DIST_TIMEOUT_S"); double s = e ? atof(e) : 120.0; return (s > 0 ? s : 120.0) * 1e3; }();
    return v;
  }
  const fl_t *round_wait(int slot, uint32_t seq) {
    volatile uint32_t *flag = &ctx->h_slots[slot].seq;
    double t_start = 0;
    for (uint64_t spins = 1;; spins++) {
      if (*flag == seq) break;
      if ((spins & 0x3ff) == 0) {  // a faulted kernel must not hang the host
        cudaError_t e = cudaStreamQuery(st);
        if (e == cudaSuccess) {
          if (*flag == seq) break;
          throw Error(VPIN_ERR_CUDA, "round result never arrived");
        }
        if (e != cudaErrorNotReady) VPIN_CUDA(e);
        if (ctx->world > 1 && (spins & 0xfffff) == 0) {
          double now = now_ms();
          if (t_start == 0) t_start = now;
          else if (now - t_start > dist_timeout_ms()) {
            dist_abort(ctx);
            throw Error(VPIN_ERR_CUDA, "a sumcheck round timed out on a distributed context (a peer rank failed?): communicator aborted");
          }
        }
      }
      __builtin_ia32_pause();
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return ctx->h_slots[slot].vals;
  }
  fl_t dot_dev(const fl_t *a, const fl_t *b, size_t n) {
    {
      ProfScope ps(ctx, PROF_DOT, (double)n, 64.0 * n, 2);
      launch_dot(a, b, n, ctx->d_small.p + 250, ctx->d_partials.p, st);
    }
    return down1(ctx->d_small.p + 250);
  }

  // ---- host commitments with the sat generators ----
  hge_t commit1(const PcGens &pc, const fl_t &x, const fl_t &blind) {  // SP/commitments.rs:79-84
    hge_t acc = hf::ge_identity();
    pc.g1->mul_acc(x, &acc);
    pc.h->mul_acc(blind, &acc);
    return acc;
  }
  hge_t commit_coeffs(const std::vector<fl_t> &c, const fl_t &blind) {  // gens_3 / gens_4 (SP/r1csproof.rs:62-72)
    hge_t acc = hf::ge_identity();
    for (size_t i = 0; i < c.size(); i++) g.sat_g[i]->mul_acc(c[i], &acc);
    g.sat_g[c.size()]->mul_acc(blind, &acc);
    return acc;
  }
  // the same two commitments with their fixed-base multiplications spread over the helper threads (inside a HostPool::Scope);
  // the group element, hence the encoding, does not depend on the order of the additions
  hge_t commit1_par(const PcGens &pc, const fl_t &x, const fl_t &blind) {
    hge_t part = hf::ge_identity(), acc = hf::ge_identity();
    HostPool::Joiner join(pool);
    pool.run(0, [&] { pc.g1->mul_acc(x, &part); });
    pc.h->mul_acc(blind, &acc);
    pool.wait(0);
    return hf::ge_add(acc, part);
  }
  hge_t commit_coeffs_par(const std::vector<fl_t> &c, const fl_t &blind) {
    const size_t n = c.size();  // 3 (quadratic round) or 4 (cubic round) coefficients + the blind
    hge_t part[2] = {hf::ge_identity(), hf::ge_identity()}, acc = hf::ge_identity();
    HostPool::Joiner join(pool);
    pool.run(0, [&] { g.sat_g[0]->mul_acc(c[0], &part[0]); g.sat_g[1]->mul_acc(c[1], &part[0]); });
    pool.run(1, [&] { g.sat_g[2]->mul_acc(c[2], &part[1]); if (n > 3) g.sat_g[3]->mul_acc(c[3], &part[1]); });
    g.sat_g[n]->mul_acc(blind, &acc);
    pool.wait(0);
    pool.wait(1);
    return hf::ge_add(hf::ge_add(acc, part[0]), part[1]);
  }

  // ---- SP/unipoly.rs:23-54 ----
  static std::vector<fl_t> unipoly_from_evals(const std::vector<fl_t> &e) {
    static const fl_t one = fl_one(), two = one + one;
    static const fl_t two_inv = fl_invert(two), six_inv = fl_invert(two + two + two);
    if (e.size() == 3) {
      fl_t c = e[0];
      fl_t a = two_inv * (e[2] - e[1] - e[1] + c);
      fl_t b = e[1] - c - a;
      return {c, b, a};
    }
    fl_t d = e[0];
    fl_t a = six_inv * (e[3] - e[2] - e[2] - e[2] + e[1] + e[1] + e[1] - e[0]);
    fl_t b = two_inv * (e[0] + e[0] - e[1] - e[1] - e[1] - e[1] - e[1] + e[2] + e[2] + e[2] + e[2] - e[3]);
    fl_t c = e[1] - d - a - b;
    return {d, c, b, a};
  }
  static fl_t unipoly_eval(const std::vector<fl_t> &c, const fl_t &r) {  // :70-78
    fl_t eval = c[0], power = r;
    for (size_t i = 1; i < c.size(); i++) { eval = eval + power * c[i]; power = power * r; }
    return eval;
  }

  // ---- SP/nizk/mod.rs ----
  KnowledgeS knowledge_prove(const PcGens &pc, const fl_t &x, const fl_t &r, Comp *C_out) {  // :28-53
    t.protocol_name("knowledge proof");
    fl_t t1 = tape.scalar("t1"), t2 = tape.scalar("t2");
    Comp C = compress_host(commit1(pc, x, r));
    t.point("C", C.data());
    Comp alpha = compress_host(commit1(pc, t1, t2));
    t.point("alpha", alpha.data());
    fl_t c = t.challenge_scalar("c");
    *C_out = C;
    return KnowledgeS{alpha, x * c + t1, r * c + t2};
  }
  EqualityS equality_prove(const PcGens &pc, const fl_t &v1, const fl_t &s1, const fl_t &v2, const fl_t &s2) {  // :90-118
    t.protocol_name("equality proof");
    fl_t r = tape.scalar("r");
    Comp C1 = compress_host(commit1(pc, v1, s1));
    t.point("C1", C1.data());
    Comp C2 = compress_host(commit1(pc, v2, s2));
    t.point("C2", C2.data());
    Comp alpha = compress_host(pc.h->mul(r));
    t.point("alpha", alpha.data());
    fl_t c = t.challenge_scalar("c");
    return EqualityS{alpha, c * (s1 - s2) + r};
  }
  ProductS product_prove(const PcGens &pc, const fl_t &x, const fl_t &rX, const fl_t &y, const fl_t &rY, const fl_t &z, const fl_t &rZ,
                         Comp *Xo, Comp *Yo, Comp *Zo) {  // :162-232
    t.protocol_name("product proof");
    fl_t b1 = tape.scalar("b1"), b2 = tape.scalar("b2"), b3 = tape.scalar("b3"), b4 = tape.scalar("b4"), b5 = tape.scalar("b5");
    Comp X = compress_host(commit1(pc, x, rX));
    t.point("X", X.data());
    Comp Y = compress_host(commit1(pc, y, rY));
    t.point("Y", Y.data());
    Comp Z = compress_host(commit1(pc, z, rZ));
    t.point("Z", Z.data());
    Comp alpha = compress_host(commit1(pc, b1, b2));
    t.point("alpha", alpha.data());
    Comp beta = compress_host(commit1(pc, b3, b4));
    t.point("beta", beta.data());
    // delta = b3 * X + b5 * h with X = x*G + rX*h  ==  (b3*x) * G + (b3*rX + b5) * h   (:202-209)
    Comp delta = compress_host(commit1(pc, b3 * x, b3 * rX + b5));
    t.point("delta", delta.data());
    fl_t c = t.challenge_scalar("c");
    ProductS p;
    p.alpha = alpha; p.beta = beta; p.delta = delta;
    p.z[0] = b1 + c * x;
    p.z[1] = b2 + c * rX;
    p.z[2] = b3 + c * y;
    p.z[3] = b4 + c * rY;
    p.z[4] = b5 + c * (rZ - rX * y);
    *Xo = X; *Yo = Y; *Zo = Z;
    return p;
  }
  // DotProductProof::prove with gens_n = gens_3 / gens_4 and gens_1 of the sat gens (:315-374). Cx is the already
  // computed commitment to x_vec under blind_x (the round's comm_poly).
  // the prover's random draws of one DotProductProof (d_vec, r_delta, r_beta: SP/nizk/mod.rs:326-333) and the commitment
  // delta = d_vec.commit(r_delta), which depends on nothing else and can therefore be computed ahead of its round
  struct DotDraw { std::vector<fl_t> d_vec; fl_t r_delta, r_beta; Comp delta; };
  DotDraw dot_draw(size_t n) {
    DotDraw d;
    d.d_vec = tape.vector("d_vec", n);
    d.r_delta = tape.scalar("r_delta");
    d.r_beta = tape.scalar("r_beta");
    return d;
  }
  DotProductProofS dotproduct_prove(const std::vector<fl_t> &x_vec, const fl_t &blind_x, const Comp &Cx, const std::vector<fl_t> &a_vec,
                                    const fl_t &y, const fl_t &blind_y, const DotDraw &draw) {
    t.protocol_name("dot product proof");
    size_t n = x_vec.size();
    const std::vector<fl_t> &d_vec = draw.d_vec;
    const fl_t &r_delta = draw.r_delta, &r_beta = draw.r_beta;
    // Cy = y G1 + blind_y h and beta = <a, d> G1 + r_beta h depend on nothing the transcript produces in between: the four
    // multiplications run side by side (called inside zk_sumcheck's HostPool::Scope), then the two encodings
    fl_t dot = fl_zero();
    for (size_t i = 0; i < n; i++) dot = dot + a_vec[i] * d_vec[i];
    hge_t p_y = hf::ge_identity(), p_by = hf::ge_identity(), p_dot = hf::ge_identity(), p_rb = hf::ge_identity();
    Comp beta;
    hge_t beta_pt;
    HostPool::Joiner join(pool);
    pool.run(0, [&] { g.sat_pc.g1->mul_acc(y, &p_y); });
    pool.run(1, [&] { g.sat_pc.h->mul_acc(blind_y, &p_by); });
    pool.run(2, [&] { g.sat_pc.g1->mul_acc(dot, &p_dot); });
    g.sat_pc.h->mul_acc(r_beta, &p_rb);
    pool.wait(2);
    beta_pt = hf::ge_add(p_dot, p_rb);
    pool.run(2, [&] { beta = compress_host(beta_pt); });
    pool.wait(0);
    pool.wait(1);
    Comp Cy = compress_host(hf::ge_add(p_y, p_by));
    pool.wait(2);
    t.point("Cx", Cx.data());
    t.point("Cy", Cy.data());
    t.scalars("a", a_vec);
    const Comp &delta = draw.delta;
    t.point("delta", delta.data());
    t.point("beta", beta.data());
    fl_t c = t.challenge_scalar("c");
    DotProductProofS p;
    p.delta = delta; p.beta = beta;
    p.z.resize(n);
    for (size_t i = 0; i < n; i++) p.z[i] = c * x_vec[i] + d_vec[i];
    p.z_delta = c * blind_x + r_delta;
    p.z_beta = c * blind_y + r_beta;
    return p;
  }

  // ---- ZK sumchecks (SP/sumcheck.rs:428-776). launch_round(q, bind, r, ctl) enqueues the fused kernel of one round
  // (bind every table with r, then evaluate `degree` sums over q thread items); launch_final(r, ctl) the last bind, which
  // delivers the final claims (`nfinal` values, returned in *finals). The kernel of round j+1 is launched as soon as
  // r_j is known, so the device works on it while the host finishes round j's sigma protocol. ----
  template <class RoundFn, class FinalFn>
  // split_eq (optional): the point tau of an eq(tau, .) factor of the summand that launch_round has split off - the kernel then
  // delivers (t(0), t(inf)) of the quadratic cofactor instead of three evaluations, and *eq_final receives eq(tau, r).
  ZkSumcheckS zk_sumcheck(const fl_t &claim, const fl_t &blind_claim, size_t num_rounds, size_t len, int degree, RoundFn launch_round,
                          FinalFn launch_final, size_t nfinal, std::vector<fl_t> *r_out, fl_t *blind_post, std::vector<fl_t> *finals,
                          const std::vector<fl_t> *split_eq = nullptr, fl_t *eq_final = nullptr) {
    std::vector<fl_t> blinds_poly = tape.vector("blinds_poly", num_rounds);
    std::vector<fl_t> blinds_evals = tape.vector("blinds_evals", num_rounds);
    // The tape is touched by nothing but the per-round DotProductProofs from here to the end of the sumcheck, so their
    // draws can be taken now in the same order, and a worker thread commits to the d_vecs (5 fixed-base multiplications
    // and an encoding per round) while this thread runs the rounds.
    std::vector<DotDraw> draws(num_rounds);
    for (size_t j = 0; j < num_rounds; j++) draws[j] = dot_draw((size_t)degree + 1);
    std::atomic<size_t> deltas_ready{0};
    std::thread delta_worker([&] {
      for (size_t j = 0; j < num_rounds; j++) {
        draws[j].delta = compress_host(commit_coeffs(draws[j].d_vec, draws[j].r_delta));
        deltas_ready.store(j + 1, std::memory_order_release);
      }
    });
    struct Joiner { std::thread &th; ~Joiner() { if (th.joinable()) th.join(); } } joiner{delta_worker};
    HostPool::Scope helpers(pool);
    fl_t claim_per_round = claim;
    Comp comm_claim_per_round = compress_host(commit1(g.sat_pc, claim_per_round, blind_claim));
    ZkSumcheckS out;
    std::vector<fl_t> r;
    uint32_t seq;
    launch_round(len >> 1, false, fl_zero(), round_ctl(0, &seq));
    std::vector<fl_t> winv;
    if (split_eq) winv = batch_invert(*split_eq);  // (while the first round's kernel runs)
    fl_t E = fl_one(), sc = claim;  // E = eq(tau_<j, r_<j); sc = claim_per_round / E
    SplitEq keep;
    for (size_t j = 0; j < num_rounds; j++) {
      fl_t ev[3];
      if (split_eq) {
        fl_t tv[2];
        memcpy(tv, round_wait((int)(j & 1), seq), sizeof(tv));
        SplitEq::evals(E, (*split_eq)[j], winv[j], sc, tv[0], tv[1], ev, &keep);
      } else {
        memcpy(ev, round_wait((int)(j & 1), seq), degree * sizeof(fl_t));
      }
      std::vector<fl_t> evals = degree == 3 ? std::vector<fl_t>{ev[0], claim_per_round - ev[0], ev[1], ev[2]}
                                            : std::vector<fl_t>{ev[0], claim_per_round - ev[0], ev[1]};
      std::vector<fl_t> poly = unipoly_from_evals(evals);
      Comp comm_poly = compress_host(commit_coeffs_par(poly, blinds_poly[j]));
      t.point("comm_poly", comm_poly.data());
      out.comm_polys.push_back(comm_poly);
      fl_t r_j = t.challenge_scalar("challenge_nextround");
      if (j + 1 < num_rounds) launch_round(len >> (j + 2), true, r_j, round_ctl((int)((j + 1) & 1), &seq));
      else launch_final(r_j, round_ctl((int)((j + 1) & 1), &seq));
      fl_t eval = unipoly_eval(poly, r_j);
      if (split_eq) { E = E * keep.l_at(r_j); sc = keep.t_at(r_j); }
      Comp comm_eval = compress_host(commit1_par(g.sat_pc, eval, blinds_evals[j]));
      t.point("comm_claim_per_round", comm_claim_per_round.data());
      t.point("comm_eval", comm_eval.data());
      std::vector<fl_t> w = t.challenge_vector("combine_two_claims_to_one", 2);
      fl_t target = w[0] * claim_per_round + w[1] * eval;
      const fl_t &blind_sc = j == 0 ? blind_claim : blinds_evals[j - 1];
      fl_t blind = w[0] * blind_sc + w[1] * blinds_evals[j];
      // the reference asserts target.commit(blind) == w0*comm_claim + w1*comm_eval (:531, :722); it holds by
      // linearity of the commitments computed above, so the variable-base MSM is not recomputed here.
      size_t deg = poly.size() - 1;
      std::vector<fl_t> a(deg + 1);
      fl_t pw = fl_one();
      for (size_t i = 0; i <= deg; i++) {
        fl_t a_sc = i == 0 ? fl_one() + fl_one() : fl_one();
        a[i] = w[0] * a_sc + w[1] * pw;
        pw = pw * r_j;
      }
      while (deltas_ready.load(std::memory_order_acquire) <= j) __builtin_ia32_pause();
      out.proofs.push_back(dotproduct_prove(poly, blinds_poly[j], comm_poly, a, target, blind, draws[j]));
      claim_per_round = eval;
      comm_claim_per_round = comm_eval;
      r.push_back(r_j);
      out.comm_evals.push_back(comm_eval);
    }
    const fl_t *fin = round_wait((int)(num_rounds & 1), seq);
    finals->assign(fin, fin + nfinal);
    if (eq_final) *eq_final = E;
    *r_out = r;
    *blind_post = blinds_evals[num_rounds - 1];
    return out;
  }

  // ---- DotProductProofLog::prove (SP/nizk/mod.rs:447-531) + BulletReductionProof::prove (SP/nizk/bullet.rs:32-132).
  // x_vec, a_vec: device vectors of length n = pc.R. The reference folds the generators every round
  // (G_L[i] = u^-1 G_L[i] + u G_R[i]); here the folded generators are never materialised: a weight vector W over the
  // ORIGINAL generators tracks the folding, so every L, R and g_hat is a fixed-base MSM over the same table.
  DotLogS dotproductlog_prove(const PcGens &pc, const LabelGens &lg, const fl_t *d_x, const fl_t &blind_x, const fl_t *d_a, const fl_t &y,
                              const fl_t &blind_y, Comp *Cy_out) {
    t.protocol_name("dot product proof (log)");
    double tb0 = now_ms();
    size_t n = pc.R, lg_n = math_log2(n);
    fl_t d = tape.scalar("d");
    fl_t r_delta = tape.scalar("r_delta");
    fl_t r_beta = tape.scalar("r_delta");  // sic (mod.rs:466)
    std::vector<fl_t> bv1 = tape.vector("blinds_vec_1", 2 * lg_n), bv2 = tape.vector("blinds_vec_2", 2 * lg_n);
    const MsmGeom geom = lg.geom;
    // Horner pass over the geom.group window sums of one row (kernels_msm.cuh), on the host: 60 doublings cost 14 us here
    // and 130-250 us in a lone GPU thread
    auto horner = [&](const fl_t *vals, size_t row) {
      ge_t w[kMsmMaxGroup];
      memcpy(w, reinterpret_cast<const uint8_t *>(vals + 8) + row * geom.group * sizeof(ge_t), geom.group * sizeof(ge_t));
      hge_t h = hf::ge_from_dev(w[geom.group - 1]);
      for (int k = geom.group - 2; k >= 0; k--) {
        for (int i = 0; i < geom.W; i++) h = hf::ge_dbl(h);
        h = hf::ge_add(h, hf::ge_from_dev(w[k]));
      }
      return h;
    };
    // Cx = x_vec.commit(blind_x, gens_n): one-row fixed-base MSM whose window sums land in the host-mapped slot; the copy of
    // a_vec for the transcript travels behind it on the same stream
    hge_t cx;
    uint32_t cx_seq = 0;
    std::vector<fl_t> a_host(n);
    {
      size_t segs1 = msm_num_segments(1, n, geom);
      DevVec<uint16_t> dg(msm_digits_count(1, n, geom), st);
      DevVec<ge_t> part(geom.group * segs1, st);
      {
        ProfScope ps(ctx, PROF_MSM_RECODE, (double)n, (double)n * (32 + 2 * geom.windows));
        launch_recode(d_x, 1, n, n, nullptr, geom, dg.p, ctx->d_counters.p, st);
      }
      {
        ProfScope ps(ctx, PROF_MSM_ACCUMULATE, (double)n, 0);
        launch_msm_accumulate(lg.table(), dg.p, 1, n, false, 0, segs1, part.p, st);
      }
      uint32_t seq = ++ctx->round_seq;
      {  // the last block of the segment-sum kernel publishes the sequence number (slot 1: slot 0 belongs to round 0 below)
        ProfScope ps(ctx, PROF_MSM_FINISH, (double)n, 0);
        launch_msm_segsum(part.p, 1, segs1, geom, reinterpret_cast<ge_t *>(ctx->d_slots[1].vals + 8), st, ctx->d_round_counters.p + 31,
                          &ctx->d_slots[1].seq, seq);
      }
      VPIN_CUDA(cudaMemcpyAsync(a_host.data(), d_a, n * sizeof(fl_t), cudaMemcpyDeviceToHost, st));
      ctx->mark();
      cx_seq = seq;
    }
    // a / b ping-pong buffers, weights over the original generators, MSM scratch for two rows (kernels_round.cu)
    DevVec<fl_t> abuf(2 * n, st), bbuf(2 * n, st), W(n, st);
    fl_t *av[2] = {abuf.p, abuf.p + n}, *bv[2] = {bbuf.p, bbuf.p + n};
    VPIN_CUDA(cudaMemcpyAsync(av[0], d_x, n * sizeof(fl_t), cudaMemcpyDeviceToDevice, st));
    VPIN_CUDA(cudaMemcpyAsync(bv[0], d_a, n * sizeof(fl_t), cudaMemcpyDeviceToDevice, st));
    launch_fill_one(W.p, n, st);
    size_t segs = msm_num_segments(2, n, geom), stride = msm_col_stride(n);
    DevVec<uint16_t> digits(msm_digits_count(2, n, geom), st);
    DevVec<ge_t> partial(2 * geom.group * segs, st);
    DotLogS out;
    int src = 0, slot = 0;
    bool fold = false;
    fl_t u = fl_zero(), u_inv = fl_zero();
    // one round = k_bullet_round (fold with the previous u, c_L / c_R, MSM digits) + the two-row fixed-base MSM whose
    // Horner kernel stores L', R' straight into the host-mapped slot
    auto launch = [&](size_t len, bool final, uint32_t *seq_out) {
      BulletRoundArgs p;
      p.a_old = av[src]; p.b_old = bv[src]; p.a_new = av[src ^ 1]; p.b_new = bv[src ^ 1];
      p.W = W.p; p.n = n; p.len = len; p.u = u; p.uinv = u_inv; p.d = d;
      p.fold = fold ? 1 : 0; p.final = final ? 1 : 0;
      p.geom = geom; p.digits = digits.p; p.stride = stride; p.nonzero = ctx->d_counters.p;
      uint32_t seq0;
      p.ctl = round_ctl(slot, &seq0);
      size_t rows = final ? 1 : 2;
      double pts = (double)rows * n;
      {
        ProfScope ps(ctx, PROF_BULLET, pts, 0);
        launch_bullet_round(p, st);
      }
      size_t sg = final ? msm_num_segments(1, n, geom) : segs;
      if (sg > segs) sg = segs;
      {
        ProfScope ps(ctx, PROF_MSM_ACCUMULATE, pts, 0);
        launch_msm_accumulate(lg.table(), digits.p, rows, n, false, 0, sg, partial.p, st);
      }
      *seq_out = ++ctx->round_seq;
      {  // window sums straight into the host-mapped slot: vals[8 ..] = rows x kMsmGroup points; the kernel's last block publishes
        ProfScope ps(ctx, PROF_MSM_FINISH, pts, 0);
        launch_msm_segsum(partial.p, rows, sg, geom, reinterpret_cast<ge_t *>(ctx->d_slots[slot].vals + 8), st, ctx->d_round_counters.p + 31,
                          &ctx->d_slots[slot].seq, *seq_out);
      }
      if (fold && !final) src ^= 1;
    };
    static_assert(8 * sizeof(fl_t) + 2 * kMsmMaxGroup * sizeof(ge_t) <= kRoundSlotVals * sizeof(fl_t), "slot too small for two rows");
    // Round 0 needs nothing from the transcript (no fold yet): its kernels are queued now and run while the host finishes Cx
    // and absorbs the n scalars of a_vec (0.3 ms of Keccak at n = 4096)
    size_t len = n;
    uint32_t seq_round0 = 0;
    const bool prelaunched = len != 1;
    if (prelaunched) launch(len, false, &seq_round0);
    cx = horner(round_wait(1, cx_seq), 0);
    pc.h->mul_acc(blind_x, &cx);
    Comp Cx = compress_host(cx);
    t.point("Cx", Cx.data());
    Comp Cy = compress_host(commit1(pc, y, blind_y));
    t.point("Cy", Cy.data());
    ctx->wait_mark();  // a_host has arrived (it travels behind the Cx kernels and ahead of round 0, which keeps running)
    t.scalars("a", a_host);
    fl_t r = t.challenge_scalar("r");  // Q = r * gens_1.G[0]
    fl_t blind_fin = blind_x + r * blind_y;
    t_bullet_pre += now_ms() - tb0;
    HostPool::Scope helpers(pool);
    for (size_t round = 0; len != 1; round++) {
      double tr0 = now_ms();
      uint32_t seq;
      if (round == 0 && prelaunched) seq = seq_round0;
      else launch(len, false, &seq);
      const fl_t *vals = round_wait(slot, seq);
      fl_t c[2] = {vals[0], vals[1]};
      double tr1 = now_ms();
      t_bullet_gpu += tr1 - tr0;
      const fl_t &blind_L = bv1[round], &blind_R = bv2[round];
      // L and R are independent (Horner pass, two fixed-base multiplications, encoding each): R on a helper thread
      hge_t LR[2];
      Comp Lc, Rc;
      HostPool::Joiner join(pool);
      pool.run(0, [&] {
        LR[1] = horner(vals, 1);
        pc.g1->mul_acc(c[1] * r, &LR[1]);
        pc.h->mul_acc(blind_R, &LR[1]);
        Rc = compress_host(LR[1]);
      });
      LR[0] = horner(vals, 0);
      pc.g1->mul_acc(c[0] * r, &LR[0]);
      pc.h->mul_acc(blind_L, &LR[0]);
      Lc = compress_host(LR[0]);
      pool.wait(0);
      t.point("L", Lc.data());
      t.point("R", Rc.data());
      u = t.challenge_scalar("u");
      u_inv = fl_invert(u);
      blind_fin = blind_fin + blind_L * u * u + blind_R * u_inv * u_inv;
      out.L_vec.push_back(Lc);
      out.R_vec.push_back(Rc);
      fold = true;
      len /= 2;
      slot ^= 1;
      t_bullet_host += now_ms() - tr1;
    }
    // last fold -> x_hat, a_hat; delta = d * g_hat + r_delta * h with g_hat = sum_j W_j G_j
    uint32_t seq;
    launch(1, true, &seq);
    const fl_t *vals = round_wait(slot, seq);
    fl_t x_hat = vals[0], a_hat = vals[1];
    fl_t y_hat = x_hat * a_hat;
    hge_t dl = horner(vals, 0);
    pc.h->mul_acc(r_delta, &dl);
    out.delta = compress_host(dl);
    t.point("delta", out.delta.data());
    out.beta = compress_host(commit1(pc, d * r, r_beta));  // d * Q + r_beta * h
    t.point("beta", out.beta.data());
    fl_t c = t.challenge_scalar("c");
    out.z1 = d + c * y_hat;
    out.z2 = a_hat * (c * blind_fin + r_beta) + r_delta;
    if (Cy_out) *Cy_out = Cy;
    return out;
  }

  // ---- PolyEvalProof::prove (SP/dense_mlpoly.rs:326-379). d_blinds: device blinds (L elements) or nullptr for zeros.
  DotLogS polyeval_prove(const PcGens &pc, const LabelGens &lg, const fl_t *d_Z, const fl_t *d_blinds, const std::vector<fl_t> &r,
                         const fl_t &Zr, const fl_t &blind_Zr, Comp *C_Zr_prime) {
    t.protocol_name("polynomial evaluation proof");
    VPIN_REQUIRE(r.size() == pc.ell, VPIN_ERR_SIZE_MISMATCH, "polyeval: point size");
    size_t l = pc.ell / 2;
    DevVec<fl_t> dL = eq_table(std::vector<fl_t>(r.begin(), r.begin() + l));
    DevVec<fl_t> dR = eq_table(std::vector<fl_t>(r.begin() + l, r.end()));
    DevVec<fl_t> LZ(pc.R, st), tmp(64 * pc.R, st);
    {
      ProfScope ps(ctx, PROF_BOUND, (double)pc.L * pc.R, 32.0 * pc.L * pc.R, 2);
      launch_bound(d_Z, dL.p, pc.L, pc.R, LZ.p, tmp.p, st);
    }
    fl_t LZ_blind = d_blinds ? dot_dev(d_blinds, dL.p, pc.L) : fl_zero();
    return dotproductlog_prove(pc, lg, LZ.p, LZ_blind, dR.p, Zr, blind_Zr, C_Zr_prime);
  }
  // DensePolynomial::evaluate (SP/dense_mlpoly.rs:249-255) of a device table
  fl_t evaluate_dev(const fl_t *d_Z, const std::vector<fl_t> &r) {
    DevVec<fl_t> chis = eq_table(r);
    return dot_dev(d_Z, chis.p, (size_t)1 << r.size());
  }

  // ---- ProductCircuitEvalProofBatched::prove (SP/product_tree.rs:259-383) with prove_cubic_batched
  // (SP/sumcheck.rs:254-424). trees[c]: the circuit's layers packed [V_0 | V_1 | ... | V_last] (V_0 = n leaves).
  // dotp: optional (left, right, weight) tables of length n/2 each, bound in place.
  struct DotpTables { fl_t *l, *r, *w; fl_t claim; };
  BatchedS batched_prove(const std::vector<fl_t *> &trees, size_t n, std::vector<DotpTables> dotp, const std::vector<fl_t> &tree_evals,
                         std::vector<fl_t> *rand_out, size_t min_sharded_len_half = kRoundDealMin) {
    BatchedS out;
    size_t num_layers = math_log2(n), nc = trees.size();
    VPIN_REQUIRE(nc + dotp.size() <= (size_t)kMaxBatched, VPIN_ERR_PROVER, "too many batched instances");
    std::vector<fl_t> claims_to_verify = tree_evals;
    std::vector<fl_t> rand;
    for (size_t layer_id = num_layers; layer_id-- > 0;) {
      size_t vlen = n >> layer_id;       // |V_layer|
      size_t off = 2 * n - 2 * vlen;     // offset of V_layer inside a packed tree
      size_t len_half = vlen / 2;
      // the product circuits share the third factor eq(rand, .): it is split off the round polynomial (k_round_cubic_batched),
      // round j reads eq(rand_{>j}, .) from the suffix tables and the table is never bound
      DevVec<fl_t> eqS = eq_suffix(rand);
      VPIN_REQUIRE(((size_t)1 << rand.size()) == len_half, VPIN_ERR_PROVER, "layer size");
      size_t num_rounds = rand.size();
      bool with_dotp = layer_id == 0 && !dotp.empty();
      size_t ninst = nc + (with_dotp ? dotp.size() : 0);
      BatchedRoundArgs args;
      memset(&args, 0, sizeof(args));
      args.nprod = (int)nc;
      for (size_t c = 0; c < nc; c++) {
        args.A[c] = trees[c] + off;
        args.B[c] = trees[c] + off + len_half;
      }
      if (with_dotp)
        for (size_t k = 0; k < dotp.size(); k++) {
          claims_to_verify.push_back(dotp[k].claim);
          args.A[nc + k] = dotp[k].l; args.B[nc + k] = dotp[k].r;
          args.Cin[nc + k] = dotp[k].w; args.Cout[nc + k] = dotp[k].w;
        }
      LayerS layer;
      std::vector<fl_t> rand_prod;
      double tables = 2.0 * nc + 1 + (with_dotp ? 3.0 * dotp.size() : 0);
      uint32_t seq = 0;
      int slot = 0;
      // Per round (index num_rounds = the final-claims kernel): sequence number of its result and, for a PRE-LAUNCHED round, the
      // mailbox tag its kernel waits for (0: the challenge travelled as a kernel parameter). Rounds [0, launched) are enqueued.
      // Pre-launch (kernels_poly.cuh ChalSlot): up to kAhead rounds are enqueued before their challenge exists; the kernel of
      // round j + 1 is already resident and polling when the host posts r_j, so neither the launch call nor the front-end
      // latency sits between two rounds. The transcript stays on the host. Only for layers that are not dealt to other ranks,
      // rounds of at most prelaunch_q() thread items (longer kernels hide their launch anyway) and never while the per-class
      // profiler brackets the launches with events.
      std::vector<uint32_t> rseq(num_rounds + 1, 0), rtag(num_rounds + 1, 0);
      size_t launched = 0, host_from = SIZE_MAX;  // host_from: first round the host evaluates itself (host tail)
      static const size_t kAhead = 3;
      ChalGuard outstanding{this, {}, {}};  // posts a dummy value to every tag still waited for if the layer unwinds
      // One proof on several GPUs (opt-in, VPIN_SHARD_SUMCHECK=1): the instances of a LARGE layer are dealt round-robin to the
      // ranks. Every layer starts from tree data that all ranks hold, and nothing outside the layer's sumcheck reads its
      // tables, so a rank only ever evaluates and binds its own instances; per round the 3 sums of every instance are
      // exchanged with one small in-place NCCL all-gather (and the final claims once per layer), then every rank continues
      // with the same transcript. Small layers are not worth the ~15 us of the collective per round and stay replicated.
      const int world = ctx->world, rank = ctx->rank;
      bool sharded = shard_sumcheck_enabled(ctx) && len_half >= min_sharded_len_half;  // (turns false when the layer un-shards)
      const size_t max_own = (ninst + world - 1) / world;
      std::vector<size_t> own;  // this rank's instances, in order
      BatchedRoundArgs own_args;
      memset(&own_args, 0, sizeof(own_args));
      if (sharded) {
        for (size_t i = rank; i < ninst; i += world) own.push_back(i);
        VPIN_REQUIRE(3 * max_own * world <= (size_t)kRoundSlotVals, VPIN_ERR_PROVER, "sharded round does not fit the slot");
        if (!ctx->d_dev_slots.p) { ctx->d_dev_slots.alloc(2, st); ctx->d_gather.alloc(kRoundSlotVals, st); }
        for (size_t k = 0; k < own.size(); k++) {
          own_args.A[k] = args.A[own[k]]; own_args.B[k] = args.B[own[k]];
          own_args.Cin[k] = args.Cin[own[k]]; own_args.Cout[k] = args.Cout[own[k]];  // (dot-product instances own their third factor)
          if (own[k] < nc) own_args.nprod = (int)k + 1;  // instances are dealt in increasing order: this rank's product circuits come first
        }
      }
      // after a sharded launch: stage this rank's values, all-gather, publish everything to the host-mapped slot
      auto exchange = [&](int per_inst) {
        const int seg = per_inst * (int)max_own;
        launch_stage_vals(ctx->d_dev_slots.p[slot].vals, per_inst * (int)own.size(), seg, ctx->d_gather.p + (size_t)rank * seg, st);
        dist_allgather_inplace(ctx, ctx->d_gather.p, (size_t)seg * sizeof(fl_t));
        seq = ++ctx->round_seq;
        launch_publish_vals(ctx->d_gather.p, seg * world, ctx->d_slots + slot, seq, st);
      };
      // gathered layout -> instance order: instance i sits at rank (i % world), position (i / world)
      auto ungather = [&](const fl_t *g, int per_inst, size_t i, int k) -> const fl_t & {
        return g[(i % world) * (per_inst * max_own) + (i / world) * per_inst + k];
      };
      // host tail (kHostTailQ): table heads on the host, [instance][A | B | C of a dot-product instance], current length `hlen`
      static const size_t host_tail_q = [] {
        const char *e = getenv("VPIN_HOST_TAIL_Q");
        long v = e ? atol(e) : (long)kHostTailQDefault;
        return (size_t)(v < 0 ? 0 : (v > 8 ? 8 : v));
      }();
      bool host_pending_bind = false;
      size_t hlen = 0;
      std::vector<std::vector<fl_t>> hT;  // 3 * ninst tables (the third one empty for product instances)
      fl_t host_r_pending = fl_zero();
      auto host_bind = [&](const fl_t &r) {  // bound_poly_var_top on every host table (SP/dense_mlpoly.rs:229-236)
        size_t half = hlen / 2;
        for (auto &T : hT)
          if (!T.empty()) {
            for (size_t i = 0; i < half; i++) T[i] = T[i] + r * (T[i + half] - T[i]);
            T.resize(half);
          }
        hlen = half;
      };
      // eq(rand[j+1..], x) for x < q on the host (the suffix table the kernels read from eqS)
      auto host_eq_rest = [&](size_t j, size_t q) {
        std::vector<fl_t> tb(1, fl_one());
        for (size_t v = num_rounds; v-- > j + 1;) {  // new variable on top of the index, as k_eq_suffix_small
          std::vector<fl_t> nx(2 * tb.size());
          for (size_t x = 0; x < tb.size(); x++) { fl_t hi = tb[x] * rand[v]; nx[x + tb.size()] = hi; nx[x] = tb[x] - hi; }
          tb.swap(nx);
        }
        VPIN_REQUIRE(tb.size() == q, VPIN_ERR_PROVER, "host tail: eq table size");
        return tb;
      };
      // the evaluations of round j from the host tables (length 2 q): what k_round_cubic_batched delivers
      auto host_eval = [&](size_t j, std::vector<fl_t> &ev_out) {
        size_t q = hlen / 2;
        std::vector<fl_t> eqr = host_eq_rest(j, q);
        for (size_t i = 0; i < ninst; i++) {
          const std::vector<fl_t> &A = hT[3 * i], &B = hT[3 * i + 1], &Cc = hT[3 * i + 2];
          fl_t a0s = fl_zero(), a1s = fl_zero(), a2s = fl_zero();
          for (size_t x = 0; x < q; x++) {
            if (i < nc) {
              a0s = a0s + eqr[x] * (A[x] * B[x]);
              a1s = a1s + eqr[x] * ((A[x + q] - A[x]) * (B[x + q] - B[x]));
            } else {
              fl_t da = A[x + q] - A[x], db = B[x + q] - B[x], dc = Cc[x + q] - Cc[x];
              fl_t a2 = A[x + q] + da, b2 = B[x + q] + db, c2 = Cc[x + q] + dc;
              a0s = a0s + (A[x] * B[x]) * Cc[x];
              a1s = a1s + (a2 * b2) * c2;
              a2s = a2s + ((a2 + da) * (b2 + db)) * (c2 + dc);
            }
          }
          ev_out[3 * i] = a0s; ev_out[3 * i + 1] = a1s; ev_out[3 * i + 2] = a2s;
        }
      };
      // instead of the kernel of round j: one copy of the table heads (4 q elements, or 2 q when no bind is pending)
      auto start_host_tail = [&](size_t j) {
        size_t q = len_half >> (j + 1);
        hlen = j > 0 ? 4 * q : 2 * q;
        FinalArgs fa;
        fa.n = 0;
        for (size_t i = 0; i < ninst; i++) {
          fa.p[fa.n++] = args.A[i];
          fa.p[fa.n++] = args.B[i];
          if (i >= nc) fa.p[fa.n++] = args.Cout[i];
        }
        VPIN_REQUIRE((size_t)fa.n * hlen <= (size_t)kTailElems, VPIN_ERR_PROVER, "host tail does not fit its buffer");
        launch_tail_copy(fa, (int)hlen, ctx->d_tail, round_ctl(slot, &seq), st);
        host_from = j;  // (the bind with r_{j-1} is delivered by `deliver` once that challenge exists)
      };
      auto finish_host_tail_copy = [&]() {  // (first use after start_host_tail: the copy has landed)
        round_wait((int)(host_from & 1), rseq[host_from]);
        hT.assign(3 * ninst, std::vector<fl_t>());
        const fl_t *src = ctx->h_tail;
        for (size_t i = 0; i < ninst; i++)
          for (int tsel = 0; tsel < (i >= nc ? 3 : 2); tsel++) {
            hT[3 * i + tsel].assign(src, src + hlen);
            src += hlen;
          }
      };
      // round j: bind with r_{j-1} (j > 0) and evaluate over q = len_half >> (j+1) thread items
      const bool pre_ok = !sharded && prelaunch_q() > 0 && !ctx->prof.on && !ctx->no_prelaunch;
      auto is_host_tail = [&](size_t q) {
        return !sharded && host_tail_q && q <= host_tail_q && (2 * ninst + (ninst - nc)) * 4 * q <= (size_t)kTailElems;
      };
      auto new_tag = [&](size_t j) {
        if (!outstanding.gate.owns_lock()) outstanding.gate = std::shared_lock<std::shared_mutex>(device_sync_gate());
        uint32_t tag = ++ctx->round_seq;
        rtag[j] = tag;
        outstanding.tags.push_back(tag);
        return ChalRef{ctx->d_chal + tag % kChalRing, ctx->d_chal_latch.p, tag, mailbox_timeout_ms() * (ctx->world > 1 ? 10u : 1u)};
      };
      // enqueues round j (launched == j). pre: the kernel takes r_{j-1} from the mailbox (r_prev is ignored)
      auto launch = [&](size_t j, const fl_t &r_prev, bool pre) {
        size_t q = len_half >> (j + 1);
        slot = (int)(j & 1);
        if (is_host_tail(q)) {
          start_host_tail(j);
          rseq[j] = seq;
          launched = num_rounds + 1;  // nothing else of this layer runs on the device
          return;
        }
        launched = j + 1;
        args.eq_rest = own_args.eq_rest = eqS.p + q;  // table k = num_rounds - 1 - j of the suffix family (2^k = q elements)
        if (sharded && j > 0 && q <= kUnshardQ) {
          // short rounds cost more in exchanges than they save: every owner broadcasts the current (4 q element) tables of its
          // instances and the layer continues replicated
          std::vector<void *> bufs;
          std::vector<size_t> bytes;
          std::vector<int> roots;
          for (size_t i = 0; i < ninst; i++) {
            fl_t *tabs[3] = {args.A[i], args.B[i], i >= nc ? args.Cout[i] : nullptr};
            for (fl_t *tb : tabs)
              if (tb) { bufs.push_back(tb); bytes.push_back(4 * q * sizeof(fl_t)); roots.push_back((int)(i % world)); }
          }
          dist_broadcast_many(ctx, bufs.data(), bytes.data(), roots.data(), (int)bufs.size());
          sharded = false;
        }
        if (sharded) {
          if (!own.empty()) {
            uint32_t dev_seq;
            RoundCtl ctl = round_ctl(slot, &dev_seq);
            ctl.slot = ctx->d_dev_slots.p + slot;  // results stay on the device until the exchange
            ProfScope ps(ctx, PROF_SC_BATCHED, (double)own.size() * q, (2.0 * own.size() + 1) * (j > 0 ? 6 : 2) * q * 32);
            launch_round_cubic_batched(own_args, (int)own.size(), q, j > 0, r_prev, ctl, st);
          }
          exchange(3);
          rseq[j] = seq;
          return;
        }
        ProfScope ps(ctx, PROF_SC_BATCHED, (double)ninst * q, tables * (j > 0 ? 6 : 2) * q * 32);
        ChalRef ch = pre ? new_tag(j) : ChalRef{nullptr, nullptr, 0, 0};
        launch_round_cubic_batched(args, (int)ninst, q, j > 0, r_prev, round_ctl(slot, &seq), st, ch);
        rseq[j] = seq;
      };
      // final claims: every table bound with the last challenge (in the kernel; the tables themselves are done with)
      FinalArgs fa;
      fa.n = 0;
      for (size_t c = 0; c < nc; c++) { fa.p[fa.n++] = args.A[c]; fa.p[fa.n++] = args.B[c]; }
      fa.p[fa.n++] = args.A[0];  // (slot of the eq claim eq(rand, r): equal to E below, not needed by the prover)
      if (with_dotp)
        for (size_t k = 0; k < dotp.size(); k++) { fa.p[fa.n++] = dotp[k].l; fa.p[fa.n++] = dotp[k].r; fa.p[fa.n++] = dotp[k].w; }
      VPIN_REQUIRE(fa.n <= kRoundSlotVals, VPIN_ERR_PROVER, "too many final claims");
      auto launch_final = [&](const fl_t &r_last, bool pre) {
        slot = (int)(num_rounds & 1);
        ProfScope ps(ctx, PROF_FINAL, (double)fa.n, 0);
        ChalRef ch = pre ? new_tag(num_rounds) : ChalRef{nullptr, nullptr, 0, 0};
        launch_round_final(fa, num_rounds > 0, r_last, round_ctl(slot, &seq), st, ch);
        rseq[num_rounds] = seq;
        launched = num_rounds + 1;
      };
      // enqueues, ahead of their challenges, the rounds before `upto` (and the final-claims kernel after the last round)
      auto launch_ahead = [&](size_t upto) {
        if (!pre_ok) return;
        while (launched < upto && launched <= num_rounds) {
          size_t jj = launched;
          if (jj == num_rounds) { launch_final(fl_zero(), true); break; }
          size_t q = len_half >> (jj + 1);
          bool tail = is_host_tail(q);
          if (!tail && q > prelaunch_q()) break;
          launch(jj, fl_zero(), !tail);
        }
      };
      // r = r_{next-1} has been drawn: hand it to round `next` (next == num_rounds: the final claims)
      auto deliver = [&](size_t next, const fl_t &r) {
        if (next >= host_from) {  // the host binds its own tables
          if (next < num_rounds) { host_pending_bind = true; host_r_pending = r; }
          return;
        }
        if (next < launched) {  // pre-launched: its kernel is waiting for the tag (or about to)
          if (rtag[next]) chal_post(rtag[next], r, &outstanding);
          return;
        }
        if (next < num_rounds) {
          launch(next, r, false);
          if (next >= host_from) { host_pending_bind = true; host_r_pending = r; }  // (the launch started the host tail)
        }
      };
      fl_t r_j = fl_zero();
      if (num_rounds) launch(0, r_j, false);
      // the first round's kernel needs neither the coefficients nor the claim: both are drawn / formed while it runs
      std::vector<fl_t> coeffs = t.challenge_vector("rand_coeffs_next_layer", claims_to_verify.size());
      fl_t e = fl_zero();
      for (size_t i = 0; i < coeffs.size(); i++) e = e + claims_to_verify[i] * coeffs[i];
      std::vector<fl_t> ev(3 * ninst);
      // split-eq bookkeeping of the product instances (SplitEq): E = eq(rand_<j, r_<j). Every product instance has the same
      // factor E l(X), so the coefficients are folded in BEFORE the split is undone: with T = sum_c coeffs[c] t_c (a quadratic)
      // sum_c coeffs[c] s_c(X) = E l(X) T(X), and one SplitEq (on T(0), T(inf) and the scaled claim SC = sum_c coeffs[c] sc_c)
      // replaces one per instance: ~36 instead of ~200 host multiplications per round. Exact field arithmetic (linearity).
      std::vector<fl_t> winv = batch_invert(rand);
      fl_t SC = fl_zero();
      for (size_t c = 0; c < nc; c++) SC = SC + claims_to_verify[c] * coeffs[c];
      SplitEq keep;
      fl_t E = fl_one();
      for (size_t j = 0; j < num_rounds; j++) {
        launch_ahead(j + 1 + kAhead);
        double tw0 = now_ms();
        if (j >= host_from) {
          if (hT.empty()) finish_host_tail_copy();
          if (host_pending_bind) { host_bind(host_r_pending); host_pending_bind = false; }
          host_eval(j, ev);
        } else {
          slot = (int)(j & 1);
          seq = rseq[j];
          const fl_t *got = round_wait(slot, seq);
          if (sharded) {
            for (size_t i = 0; i < ninst; i++)
              for (int k = 0; k < 3; k++) ev[3 * i + k] = ungather(got, 3, i, k);
          } else {
            memcpy(ev.data(), got, 3 * ninst * sizeof(fl_t));
          }
        }
        double tw1 = now_ms();
        t_b_wait += tw1 - tw0;
        n_b_rounds++;
        if ((len_half >> (j + 1)) <= 64) { t_b_small_wait += tw1 - tw0; n_b_small++; }
        fl_t c0 = fl_zero(), c2 = fl_zero(), c3 = fl_zero();
        if (nc) {  // (T(0), T(inf)) of the combined cofactor -> the combined evaluations at 0, 2, 3 the reference computes
          fl_t T0 = fl_zero(), Ainf = fl_zero(), out3[3];
          for (size_t c = 0; c < nc; c++) {
            T0 = T0 + ev[3 * c] * coeffs[c];
            Ainf = Ainf + ev[3 * c + 1] * coeffs[c];
          }
          SplitEq::evals(E, rand[j], winv[j], SC, T0, Ainf, out3, &keep);
          c0 = out3[0]; c2 = out3[1]; c3 = out3[2];
        }
        for (size_t i = nc; i < ninst; i++) {
          c0 = c0 + ev[3 * i] * coeffs[i];
          c2 = c2 + ev[3 * i + 1] * coeffs[i];
          c3 = c3 + ev[3 * i + 2] * coeffs[i];
        }
        std::vector<fl_t> poly = unipoly_from_evals({c0, e - c0, c2, c3});
        // UniPoly::append_to_transcript (SP/unipoly.rs:111-119)
        t.message("poly", "UniPoly_begin");
        for (auto &c : poly) t.scalar("coeff", c);
        t.message("poly", "UniPoly_end");
        r_j = t.challenge_scalar("challenge_nextround");
        rand_prod.push_back(r_j);
        double tw2 = now_ms();
        t_b_host += tw2 - tw1;
        deliver(j + 1, r_j);
        t_b_launch += now_ms() - tw2;
        if (nc) { SC = keep.t_at(r_j); E = E * keep.l_at(r_j); }
        e = unipoly_eval(poly, r_j);
        layer.polys.push_back({poly[0], poly[2], poly[3]});  // CompressedUniPoly (SP/unipoly.rs:80-87)
      }
      std::vector<fl_t> fin(fa.n);
      slot = (int)(num_rounds & 1);
      if (host_from != SIZE_MAX) {  // the last bind on the host; fin layout as below: (left, right) per circuit | eq claim slot | (l, r, w) per dot product
        if (hT.empty()) finish_host_tail_copy();
        if (num_rounds > 0) host_bind(r_j);
        VPIN_REQUIRE(hlen == 1, VPIN_ERR_PROVER, "host tail: table length");
        for (size_t c = 0; c < nc; c++) { fin[2 * c] = hT[3 * c][0]; fin[2 * c + 1] = hT[3 * c + 1][0]; }
        fin[2 * nc] = fl_zero();
        if (with_dotp)
          for (size_t k = 0; k < dotp.size(); k++)
            for (int x = 0; x < 3; x++) fin[2 * nc + 1 + 3 * k + x] = hT[3 * (nc + k) + x][0];
      } else if (sharded) {  // three final claims per own instance (left, right, third factor), exchanged like the round sums
        FinalArgs fo;
        fo.n = 0;
        for (size_t k = 0; k < own.size(); k++) {
          fo.p[fo.n++] = own_args.A[k];
          fo.p[fo.n++] = own_args.B[k];
          fo.p[fo.n++] = own[k] < nc ? own_args.A[k] : own_args.Cin[k];
        }
        if (fo.n) {
          uint32_t dev_seq;
          RoundCtl ctl = round_ctl(slot, &dev_seq);
          ctl.slot = ctx->d_dev_slots.p + slot;
          launch_round_final(fo, num_rounds > 0, r_j, ctl, st);
        }
        exchange(3);
        const fl_t *got = round_wait(slot, seq);
        for (size_t c = 0; c < nc; c++) { fin[2 * c] = ungather(got, 3, c, 0); fin[2 * c + 1] = ungather(got, 3, c, 1); }
        fin[2 * nc] = ungather(got, 3, 0, 2);
        if (with_dotp)
          for (size_t k = 0; k < dotp.size(); k++)
            for (int x = 0; x < 3; x++) fin[2 * nc + 1 + 3 * k + x] = ungather(got, 3, nc + k, x);
      } else {
        if (launched <= num_rounds) launch_final(r_j, false);
        memcpy(fin.data(), round_wait((int)(num_rounds & 1), rseq[num_rounds]), fa.n * sizeof(fl_t));
      }
      for (size_t c = 0; c < nc; c++) { layer.left.push_back(fin[2 * c]); layer.right.push_back(fin[2 * c + 1]); }
      for (size_t c = 0; c < nc; c++) {
        t.scalar("claim_prod_left", layer.left[c]);
        t.scalar("claim_prod_right", layer.right[c]);
      }
      if (with_dotp) {
        for (size_t k = 0; k < dotp.size(); k++) {
          fl_t l = fin[2 * nc + 1 + 3 * k], r = fin[2 * nc + 2 + 3 * k], w = fin[2 * nc + 3 + 3 * k];
          t.scalar("claim_dotp_left", l);
          t.scalar("claim_dotp_right", r);
          t.scalar("claim_dotp_weight", w);
          out.dotp[0].push_back(l); out.dotp[1].push_back(r); out.dotp[2].push_back(w);
        }
      }
      fl_t r_layer = t.challenge_scalar("challenge_r_layer");
      claims_to_verify.clear();
      for (size_t c = 0; c < nc; c++) claims_to_verify.push_back(layer.left[c] + r_layer * (layer.right[c] - layer.left[c]));
      std::vector<fl_t> ext = {r_layer};
      ext.insert(ext.end(), rand_prod.begin(), rand_prod.end());
      rand = ext;
      out.layers.push_back(std::move(layer));
    }
    *rand_out = rand;
    return out;
  }
};

}  // namespace

// builds the packed product tree of `n` leaves already stored at tree[0..n) (SP/product_tree.rs:18-56): layers of length
// n, n/2, ..., 2 (the two factors of the root), 2n - 2 elements in all
void build_trees(Ctx *ctx, const std::vector<fl_t *> &trees, size_t n, cudaStream_t st) {
  VPIN_REQUIRE(trees.size() <= 16, VPIN_ERR_PROVER, "too many trees in one batch");
  if (trees.empty() || n <= 2) return;
  TreeBatch b;
  b.n = (int)trees.size();
  for (int k = 0; k < b.n; k++) b.p[k] = trees[k];
  // every layer reads 2 x 32 B and writes 32 B per product: 48 B x (n + n/2 + ... + 4) per tree
  ProfScope ps(ctx, PROF_TREE, (double)b.n * (double)n, 48.0 * 2.0 * (double)n * b.n, 8);
  launch_build_trees(b, n, st);
}
void build_tree(Ctx *ctx, fl_t *tree, size_t n, cudaStream_t st) { build_trees(ctx, std::vector<fl_t *>{tree}, n, st); }

bool instance_is_sat(Ctx *ctx, const Instance &inst, const uint8_t *vars32, uint64_t n_vars, const uint8_t *inputs32, uint64_t n_inputs) {
  cudaStream_t st = ctx->st;
  size_t nz = 2 * inst.num_vars;
  std::vector<fl_t> z(nz, fl_zero());
  for (size_t i = 0; i < n_vars; i++) VPIN_REQUIRE(fl_from_bytes(vars32 + 32 * i, &z[i]), VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
  z[inst.num_vars] = fl_one();
  for (size_t i = 0; i < n_inputs; i++) VPIN_REQUIRE(fl_from_bytes(inputs32 + 32 * i, &z[inst.num_vars + 1 + i]), VPIN_ERR_INVALID_SCALAR, "InvalidScalar");
  DevVec<fl_t> dz(nz, st), out(3 * inst.num_cons, st);
  dz.upload(z.data(), nz);
  for (int k = 0; k < 3; k++) launch_spmv_csr(csr_of(inst.M[k], inst.num_cons), dz.p, out.p + k * inst.num_cons, st);
  std::vector<fl_t> h(3 * inst.num_cons);
  out.download(h.data(), h.size());
  ctx->sync();
  for (size_t i = 0; i < inst.num_cons; i++)
    if (!fl_eq(h[i] * h[inst.num_cons + i], h[2 * inst.num_cons + i])) return false;
  return true;
}

// VP/commit_test.rs:59-334 (my_lib_prove + my_R1CSProof_prove) and SP/sparse_mlpoly.rs:1466-1533 (SPARK)
std::vector<uint8_t> snark_prove(Ctx *ctx, const Instance &inst, const Decomm &dec, const Witness &wit, const std::vector<fl_t> &inputs,
                                 const SnarkGens &g, const uint8_t *label, size_t label_len, const fl_t &tape_seed) {
  cudaStream_t st = ctx->st;
  ctx->phases.clear();
  double t_prove = now_ms(), t0;
  auto phase = [&](const char *name, double start) { ctx->sync(); ctx->phases.push_back({name, now_ms() - start}); };
  MerlinTranscript t(label, label_len);
  ProverTape tape("proof", 5, tape_seed);
  Prover P{ctx, st, t, tape, g, host_pool_of(ctx)};
  // a mailbox time-out of an earlier, failed call must not fail this one (the abort word of the latch is sticky)
  VPIN_CUDA(cudaMemsetAsync(&ctx->d_chal_latch.p->abort, 0, sizeof(uint32_t), st));
  const PcGens &spc = g.sat_pc;
  size_t num_vars = inst.num_vars, num_cons = inst.num_cons, num_inputs = inputs.size();
  VPIN_REQUIRE(wit.n_vars == num_vars && num_inputs < num_vars, VPIN_ERR_SIZE_MISMATCH, "witness size");
  VPIN_REQUIRE(spc.L * spc.R == num_vars && wit.comm.size() == 32 * spc.L, VPIN_ERR_SIZE_MISMATCH, "gens do not match the witness");

  t.protocol_name("Spartan SNARK proof");  // VP/commit_test.rs:75 (no comm append, unlike SP/lib.rs:377)
  double t_sat = now_ms();
  t.protocol_name("R1CS proof");           // :148 (no input append, unlike SP/r1csproof.rs:175)
  t0 = now_ms();
  // comm_vars.append_to_transcript (SP/dense_mlpoly.rs:305-313)
  t.message("poly_commitment", "poly_commitment_begin");
  for (size_t i = 0; i < spc.L; i++) t.point("poly_commitment_share", wit.comm.data() + 32 * i);
  t.message("poly_commitment", "poly_commitment_end");
  phase("polycommit", t0);

  t0 = now_ms();
  // z = vars || 1 || inputs || 0...   (:162-170)
  size_t zlen = 2 * num_vars;
  DevVec<fl_t> z(zlen, st);
  z.zero();
  VPIN_CUDA(cudaMemcpyAsync(z.p, wit.d_vars.p, num_vars * sizeof(fl_t), cudaMemcpyDeviceToDevice, st));
  {
    std::vector<fl_t> one_in(1 + num_inputs);
    one_in[0] = fl_one();
    for (size_t i = 0; i < num_inputs; i++) one_in[1 + i] = inputs[i];
    VPIN_CUDA(cudaMemcpyAsync(z.p + num_vars, one_in.data(), one_in.size() * sizeof(fl_t), cudaMemcpyHostToDevice, st));
    ctx->sync();
  }
  size_t num_rounds_x = math_log2(num_cons), num_rounds_y = math_log2(zlen);
  std::vector<fl_t> tau = t.challenge_vector("challenge_tau", num_rounds_x);
  DevVec<fl_t> tau_suffix = P.eq_suffix(tau);  // eq(tau_{>j}, .) of every round: the eq factor is split off the round polynomial
  DevVec<fl_t> ABCz(3 * num_cons, st);
  for (int k = 0; k < 3; k++) {
    ProfScope ps(ctx, PROF_SPMV, (double)inst.M[k].nnz, 68.0 * inst.M[k].nnz + 36.0 * num_cons);
    launch_spmv_csr(csr_of(inst.M[k], num_cons), z.p, ABCz.p + k * num_cons, st);
  }
  fl_t *Az = ABCz.p, *Bz = ABCz.p + num_cons, *Cz = ABCz.p + 2 * num_cons;
  // phase 1: sum_x eq(tau,x) (Az Bz - Cz) = 0   (SP/r1csproof.rs:94-127)
  std::vector<fl_t> rx;
  fl_t blind_claim_postsc1;
  std::vector<fl_t> fin1;
  fl_t tau_claim;
  ZkSumcheckS sc1 = P.zk_sumcheck(
      fl_zero(), fl_zero(), num_rounds_x, num_cons, 3,
      [&](size_t q, bool bind, const fl_t &r, const RoundCtl &c) {
        // algorithmic bytes of SURVEY.md 8d: 4 tables, each read once and written at half length (the split-off eq table moves less)
        ProfScope ps(ctx, PROF_SC_CUBIC, (double)q, (bind ? 4 * 192.0 : 4 * 64.0) * q);
        launch_round_r1cs_split(tau_suffix.p + q, Az, Bz, Cz, q, bind, r, c, st);
      },
      [&](const fl_t &r, const RoundCtl &c) {
        FinalArgs fa;
        fa.p[0] = Az; fa.p[1] = Bz; fa.p[2] = Cz;
        fa.n = 3;
        launch_round_final(fa, true, r, c, st);
      },
      3, &rx, &blind_claim_postsc1, &fin1, &tau, &tau_claim);
  fl_t Az_claim = fin1[0], Bz_claim = fin1[1], Cz_claim = fin1[2];
  phase("prove_sc_phase_one", t0);

  fl_t Az_blind = tape.scalar("Az_blind"), Bz_blind = tape.scalar("Bz_blind"), Cz_blind = tape.scalar("Cz_blind"),
       prod_Az_Bz_blind = tape.scalar("prod_Az_Bz_blind");
  Comp comm_Cz_claim, comm_Az_claim, comm_Bz_claim, comm_prod_Az_Bz_claims;
  KnowledgeS pok_Cz_claim = P.knowledge_prove(spc, Cz_claim, Cz_blind, &comm_Cz_claim);
  ProductS proof_prod = P.product_prove(spc, Az_claim, Az_blind, Bz_claim, Bz_blind, Az_claim * Bz_claim, prod_Az_Bz_blind, &comm_Az_claim,
                                        &comm_Bz_claim, &comm_prod_Az_Bz_claims);
  t.point("comm_Az_claim", comm_Az_claim.data());
  t.point("comm_Bz_claim", comm_Bz_claim.data());
  t.point("comm_Cz_claim", comm_Cz_claim.data());
  t.point("comm_prod_Az_Bz_claims", comm_prod_Az_Bz_claims.data());
  fl_t blind_expected_claim_postsc1 = tau_claim * (prod_Az_Bz_blind - Cz_blind);
  fl_t claim_post_phase1 = (Az_claim * Bz_claim - Cz_claim) * tau_claim;
  EqualityS proof_eq_sc_phase1 = P.equality_prove(spc, claim_post_phase1, blind_expected_claim_postsc1, claim_post_phase1, blind_claim_postsc1);

  // ---- early start of SPARK's non-deterministic witness. The row half of the derefs (mem_rx[row_addr], :267-276) depends on
  // rx only, which is final here; its Hyrax rows (3/8 of the commitment's MSM work, no transcript dependency) CAN run on a side
  // stream under the second sumcheck and the witness-polynomial evaluation proof (see derefs_early below).
  size_t N = dec.N, M = dec.M;
  const size_t num_rounds_y_ = math_log2(2 * num_vars);
  std::vector<fl_t> rx_ext = rx;  // equalize (:1448-1464)
  if (rx.size() < num_rounds_y_) { rx_ext.assign(num_rounds_y_ - rx.size(), fl_zero()); rx_ext.insert(rx_ext.end(), rx.begin(), rx.end()); }
  VPIN_REQUIRE(((size_t)1 << rx_ext.size()) == M, VPIN_ERR_SIZE_MISMATCH, "memory size");
  DevVec<fl_t> mem_rx = P.eq_table(rx_ext);
  // workspace slab: derefs 8N | mem trees 8M | ops trees 24N | dot-product clones 9N
  struct WorkspaceBusy {  // keeps the out-of-memory hook (capi.cu) away from the slab until this proof is done with it
    Ctx *c;
    explicit WorkspaceBusy(Ctx *c_) : c(c_) { c->workspace_busy = true; }
    ~WorkspaceBusy() { c->workspace_busy = false; }
  } workspace_busy(ctx);
  fl_t *ws = ctx->workspace_reserve(41 * N + 8 * M);
  struct { fl_t *p; } derefs{ws}, mem_trees{ws + 8 * N}, ops_trees{ws + 8 * N + 8 * M}, dotp_tables{ws + 32 * N + 8 * M};
  VPIN_CUDA(cudaMemsetAsync(derefs.p, 0, 8 * N * sizeof(fl_t), st));
  for (int k = 0; k < 3; k++) {
    ProfScope ps(ctx, PROF_GATHER, 1.0 * N, 1.0 * N * 68, 1);
    launch_gather(dec.row_addr[k].p, mem_rx.p, N, derefs.p + (size_t)k * N, st);
  }
  const PcGens &dpc = g.derefs_pc;
  VPIN_REQUIRE(dpc.L * dpc.R == 8 * N, VPIN_ERR_SIZE_MISMATCH, "derefs gens");
  DevVec<uint8_t> derefs_comm(32 * dpc.L, st);
  const size_t derefs_row_rows = dpc.L / 8 * 3;  // the Hyrax rows that hold row A | row B | row C
  // ON by default (VPIN_DEREFS_EARLY=0 turns it off). The side stream has the device's lowest priority and its MSM kernel is capped
  // at three resident blocks per SM (kernels_msm.cuh): without both - first measurement of round 2 - the MSM's resident blocks
  // (256 us each, six per SM, no registers left) kept the second sumcheck's round kernels waiting, phase two 2.1 -> 5.8 ms, a net
  // loss. With them, CNN A on a B200: phase two 1.8 -> 2.9 ms, evaluation proof 1.6 -> 2.6 ms, derefs commitment 8.7 -> 4.8 ms,
  // SNARK::prove 34.2 -> 31.9 ms (profiles/r2_derefs_early.log). Single GPU only (a communicator's collectives stay on one
  // stream) and not for the largest shapes (the side stream's own block cache would come out of an HBM that is planned to the
  // last gigabyte).
  const char *early_env = getenv("VPIN_DEREFS_EARLY");
  const bool derefs_early = !(early_env && atoi(early_env) == 0) && ctx->world == 1 && !ctx->prof.on && dpc.L % 8 == 0 &&
                            derefs_row_rows >= 48 && N <= ((size_t)1 << 23);
  if (derefs_early) {
    SideScope side(ctx);
    hyrax_rows(ctx, *g.eval_label, derefs.p, derefs_row_rows, dpc.R, dpc.R, nullptr, 0, nullptr, derefs_comm.p);
  }

  t0 = now_ms();
  fl_t r_A = t.challenge_scalar("challenege_Az"), r_B = t.challenge_scalar("challenege_Bz"), r_C = t.challenge_scalar("challenege_Cz");
  fl_t claim_phase2 = r_A * Az_claim + r_B * Bz_claim + r_C * Cz_claim;
  fl_t blind_claim_phase2 = r_A * Az_blind + r_B * Bz_blind + r_C * Cz_blind;
  // evals_ABC = rA * A^T eq(rx) + rB * B^T eq(rx) + rC * C^T eq(rx)   (:257-268)
  DevVec<fl_t> evals_ABC(zlen, st);
  {
    DevVec<fl_t> evals_rx = P.eq_table(rx);
    fl_t rr[3] = {r_A, r_B, r_C};
    const fl_t *d_rr = P.up(rr, 3);
    for (int k = 0; k < 3; k++) {
      ProfScope ps(ctx, PROF_SPMV_T, (double)inst.M[k].nnz, 68.0 * inst.M[k].nnz + 36.0 * zlen, 3);
      launch_spmv_csc_scaled(csc_of(inst.M[k], zlen), evals_rx.p, d_rr + k, k != 0, evals_ABC.p, ctx->d_partials.p, ctx->d_partials.n, st);
    }
    ctx->sync();
  }
  std::vector<fl_t> ry;
  fl_t blind_claim_postsc2;
  std::vector<fl_t> fin2;
  ZkSumcheckS sc2 = P.zk_sumcheck(
      claim_phase2, blind_claim_phase2, num_rounds_y, zlen, 2,
      [&](size_t q, bool bind, const fl_t &r, const RoundCtl &c) {
        ProfScope ps(ctx, PROF_SC_QUAD, (double)q, (bind ? 2 * 192.0 : 2 * 64.0) * q);
        launch_round_quad(z.p, evals_ABC.p, q, bind, r, c, st);
      },
      [&](const fl_t &r, const RoundCtl &c) {
        FinalArgs fa;
        fa.p[0] = z.p; fa.p[1] = evals_ABC.p;
        fa.n = 2;
        launch_round_final(fa, true, r, c, st);
      },
      2, &ry, &blind_claim_postsc2, &fin2);
  fl_t claims_phase2[2] = {fin2[0], fin2[1]};
  phase("prove_sc_phase_two", t0);

  t0 = now_ms();
  std::vector<fl_t> ry1(ry.begin() + 1, ry.end());
  fl_t eval_vars_at_ry = P.evaluate_dev(wit.d_vars.p, ry1);
  fl_t blind_eval = tape.scalar("blind_eval");
  Comp comm_vars_at_ry;
  DotLogS proof_eval_vars_at_ry = P.polyeval_prove(spc, *g.sat_label, wit.d_vars.p, wit.d_blinds.p, ry1, eval_vars_at_ry, blind_eval, &comm_vars_at_ry);
  phase("polyeval", t0);
  fl_t blind_eval_Z_at_ry = (fl_one() - ry[0]) * blind_eval;
  fl_t blind_expected_claim_postsc2 = claims_phase2[1] * blind_eval_Z_at_ry;
  fl_t claim_post_phase2 = claims_phase2[0] * claims_phase2[1];
  EqualityS proof_eq_sc_phase2 = P.equality_prove(spc, claim_post_phase2, blind_expected_claim_postsc2, claim_post_phase2, blind_claim_postsc2);
  phase("R1CSProof::prove", t_sat);

  // inst.evaluate(rx, ry) (SP/r1csinstance.rs:304-307) and the three claims (VP/commit_test.rs:102-108)
  t0 = now_ms();
  fl_t inst_evals[3];
  {
    DevVec<fl_t> trx = P.eq_table(rx), try_ = P.eq_table(ry);
    for (int k = 0; k < 3; k++) {
      launch_sparse_eval(inst.M[k].coo_row.p, inst.M[k].coo_col.p, inst.M[k].coo_val.p, inst.M[k].nnz, trx.p, try_.p, ctx->d_small.p + 248,
                         ctx->d_partials.p, st);
      inst_evals[k] = P.down1(ctx->d_small.p + 248);
    }
  }
  t.scalar("Ar_claim", inst_evals[0]);
  t.scalar("Br_claim", inst_evals[1]);
  t.scalar("Cr_claim", inst_evals[2]);
  phase("eval_sparse_polys", t0);

  // ---------------- R1CSEvalProof::prove -> SparseMatPolyEvalProof::prove (SP/sparse_mlpoly.rs:1466-1533) ----------------
  double t_eval = now_ms();
  t.protocol_name("Sparse polynomial evaluation proof");
  std::vector<fl_t> ry_ext = ry;  // equalize (:1448-1464); rx_ext, mem_rx and the row derefs exist since the end of phase one
  if (rx.size() > ry.size()) { ry_ext.assign(rx.size() - ry.size(), fl_zero()); ry_ext.insert(ry_ext.end(), ry.begin(), ry.end()); }
  VPIN_REQUIRE(ry.size() == num_rounds_y_ && ((size_t)1 << ry_ext.size()) == M, VPIN_ERR_SIZE_MISMATCH, "memory size");
  DevVec<fl_t> mem_ry = P.eq_table(ry_ext);
  // derefs (:525-530, :267-282) merged as row A,B,C | col A,B,C | 0 | 0 (:61)
  for (int k = 0; k < 3; k++) {
    ProfScope ps(ctx, PROF_GATHER, 1.0 * N, 1.0 * N * 68, 1);
    launch_gather(dec.col_addr[k].p, mem_ry.p, N, derefs.p + (size_t)(3 + k) * N, st);
  }
  t0 = now_ms();
  std::vector<Comp> comm_derefs(dpc.L);
  {
    const size_t r0 = derefs_early ? derefs_row_rows : 0;  // (the rows before r0 are being committed, or are done, on the side stream)
    hyrax_rows(ctx, *g.eval_label, derefs.p + r0 * dpc.R, dpc.L - r0, dpc.R, dpc.R, nullptr, 0, nullptr, derefs_comm.p + 32 * r0);
    if (derefs_early) SideScope::join(ctx);
    derefs_comm.download((uint8_t *)comm_derefs.data(), 32 * dpc.L);
    ctx->sync();
  }
  // DerefsCommitment::append_to_transcript (:216-222)
  t.message("derefs_commitment", "begin_derefs_commitment");
  t.message("comm_poly_row_col_ops_val", "poly_commitment_begin");
  for (auto &c : comm_derefs) t.point("poly_commitment_share", c.data());
  t.message("comm_poly_row_col_ops_val", "poly_commitment_end");
  t.message("derefs_commitment", "end_derefs_commitment");
  phase("commit_nondet_witness", t0);

  std::vector<fl_t> r_mem_check = t.challenge_vector("challenge_r_hash", 2);
  t0 = now_ms();
  // hash layers + product trees (:547-671). Packed trees: 2 for init/audit per side (M leaves), 6 per side for ops (N leaves)
  const fl_t *d_gt = P.up(r_mem_check);
  phase("network_alloc", t0);
  fl_t *row_init = mem_trees.p, *row_audit = mem_trees.p + 2 * M, *col_init = mem_trees.p + 4 * M, *col_audit = mem_trees.p + 6 * M;
  // ops order of the batched proof: row read A,B,C | row write A,B,C | col read A,B,C | col write A,B,C  (:1173-1187)
  std::vector<fl_t *> ops_ptr(12), mem_ptr = {row_init, row_audit, col_init, col_audit};
  for (int i = 0; i < 12; i++) ops_ptr[i] = ops_trees.p + (size_t)i * 2 * N;
  // several GPUs, large circuit sets (>= kDealBuildMin leaves): circuit i of a set belongs to rank i mod world (batched_prove
  // deals instance i to the same rank); everything of a circuit below its short upper layers then exists on its owner only
  const int world = ctx->world, rank = ctx->rank;
  const bool deal_mem = shard_sumcheck_enabled(ctx) && M >= kDealBuildMin, deal_ops = shard_sumcheck_enabled(ctx) && N >= kDealBuildMin;
  auto owns = [&](bool dealt, size_t i) { return !dealt || (int)(i % world) == rank; };
  {
    ProfScope ps(ctx, PROF_HASH, 2.0 * M, 2.0 * M * 100, 2);
    if (owns(deal_mem, 0) || owns(deal_mem, 1)) launch_hash_mem(mem_rx.p, dec.row_audit_ts.p, M, d_gt, row_init, row_audit, st);
    if (owns(deal_mem, 2) || owns(deal_mem, 3)) launch_hash_mem(mem_ry.p, dec.col_audit_ts.p, M, d_gt, col_init, col_audit, st);
  }
  for (int k = 0; k < 3; k++) {
    ProfScope ps(ctx, PROF_HASH, 2.0 * N, 2.0 * N * 104, 2);
    if (owns(deal_ops, k) || owns(deal_ops, 3 + k))
      launch_hash_ops(dec.row_addr[k].p, derefs.p + (size_t)k * N, dec.row_read_ts[k].p, N, d_gt, ops_ptr[k], ops_ptr[3 + k], st);
    if (owns(deal_ops, 6 + k) || owns(deal_ops, 9 + k))
      launch_hash_ops(dec.col_addr[k].p, derefs.p + (size_t)(3 + k) * N, dec.col_read_ts[k].p, N, d_gt, ops_ptr[6 + k], ops_ptr[9 + k], st);
  }
  phase("network_hash", t0);
  auto build_dealt = [&](const std::vector<fl_t *> &ptrs, size_t n, bool dealt) {
    std::vector<fl_t *> mine;
    for (size_t i = 0; i < ptrs.size(); i++)
      if (owns(dealt, i)) mine.push_back(ptrs[i]);
    build_trees(ctx, mine, n, st);
    if (!dealt) return;
    // the layers of at most kShardMinLayer elements (the tail of a packed tree) are proved replicated: every owner broadcasts its own
    std::vector<void *> bufs;
    std::vector<size_t> bytes;
    std::vector<int> roots;
    for (size_t i = 0; i < ptrs.size(); i++) {
      bufs.push_back(ptrs[i] + 2 * n - 2 * kShardMinLayer);
      bytes.push_back((2 * kShardMinLayer - 2) * sizeof(fl_t));
      roots.push_back((int)(i % world));
    }
    dist_broadcast_many(ctx, bufs.data(), bytes.data(), roots.data(), (int)bufs.size());
  };
  build_dealt(mem_ptr, M, deal_mem);
  build_dealt(ops_ptr, N, deal_ops);
  phase("build_layered_network", t0);

  t0 = now_ms();
  t.protocol_name("Sparse polynomial evaluation proof");     // PolyEvalNetworkProof (:1333, :1345)
  t.protocol_name("Sparse polynomial product layer proof");  // ProductLayerProof::prove (:1057)
  // circuit outputs = product of the two elements of the last layer (SP/product_tree.rs:58-63)
  // (one launch gathers the two factors of every root into the host-mapped slot)
  auto tree_evals = [&](const std::vector<fl_t *> &ptrs, size_t n) {
    FinalArgs fa;
    fa.n = 0;
    for (size_t i = 0; i < ptrs.size(); i++) { fa.p[fa.n++] = ptrs[i] + 2 * n - 4; fa.p[fa.n++] = ptrs[i] + 2 * n - 3; }
    VPIN_REQUIRE(fa.n <= kRoundSlotVals, VPIN_ERR_PROVER, "too many circuits");
    uint32_t seq;
    RoundCtl ctl = P.round_ctl(0, &seq);
    launch_round_final(fa, false, fl_zero(), ctl, st);
    const fl_t *last = P.round_wait(0, seq);
    std::vector<fl_t> ev(ptrs.size());
    for (size_t i = 0; i < ptrs.size(); i++) ev[i] = last[2 * i] * last[2 * i + 1];
    return ev;
  };
  std::vector<fl_t> mem_evals = tree_evals(mem_ptr, M), ops_evals = tree_evals(ops_ptr, N);
  struct MemClaims { fl_t init; std::vector<fl_t> read, write; fl_t audit; } er, ec;
  er.init = mem_evals[0]; er.audit = mem_evals[1]; ec.init = mem_evals[2]; ec.audit = mem_evals[3];
  er.read.assign(ops_evals.begin(), ops_evals.begin() + 3); er.write.assign(ops_evals.begin() + 3, ops_evals.begin() + 6);
  ec.read.assign(ops_evals.begin() + 6, ops_evals.begin() + 9); ec.write.assign(ops_evals.begin() + 9, ops_evals.begin() + 12);
  auto subset_check = [&](const MemClaims &m) {  // :1068-1073
    fl_t ws = fl_one(), rs = fl_one();
    for (auto &x : m.write) ws = ws * x;
    for (auto &x : m.read) rs = rs * x;
    VPIN_REQUIRE(fl_eq(m.init * ws, rs * m.audit), VPIN_ERR_PROVER, "memory check failed");
  };
  subset_check(er);
  t.scalar("claim_row_eval_init", er.init);
  t.scalars("claim_row_eval_read", er.read);
  t.scalars("claim_row_eval_write", er.write);
  t.scalar("claim_row_eval_audit", er.audit);
  subset_check(ec);
  t.scalar("claim_col_eval_init", ec.init);
  t.scalars("claim_col_eval_read", ec.read);
  t.scalars("claim_col_eval_write", ec.write);
  t.scalar("claim_col_eval_audit", ec.audit);
  // dot-product circuits on clones (row_ops_val, col_ops_val, val), split in halves (:1105-1130)
  std::vector<Prover::DotpTables> dotp(6);
  std::vector<fl_t> eval_dotp_left(3), eval_dotp_right(3);
  for (int k = 0; k < 3; k++) {
    fl_t *l = dotp_tables.p + (size_t)(3 * k) * N, *r = l + N, *w = r + N;
    VPIN_CUDA(cudaMemcpyAsync(l, derefs.p + (size_t)k * N, N * sizeof(fl_t), cudaMemcpyDeviceToDevice, st));
    VPIN_CUDA(cudaMemcpyAsync(r, derefs.p + (size_t)(3 + k) * N, N * sizeof(fl_t), cudaMemcpyDeviceToDevice, st));
    VPIN_CUDA(cudaMemcpyAsync(w, dec.val(k), N * sizeof(fl_t), cudaMemcpyDeviceToDevice, st));
    for (int hsel = 0; hsel < 2; hsel++) {
      size_t o = hsel * (N / 2);
      launch_dot3(l + o, r + o, w + o, N / 2, ctx->d_small.p + 249, ctx->d_partials.p, st);
      fl_t e = P.down1(ctx->d_small.p + 249);
      dotp[2 * k + hsel] = Prover::DotpTables{l + o, r + o, w + o, e};
      (hsel == 0 ? eval_dotp_left : eval_dotp_right)[k] = e;
    }
    t.scalar("claim_eval_dotp_left", eval_dotp_left[k]);
    t.scalar("claim_eval_dotp_right", eval_dotp_right[k]);
    VPIN_REQUIRE(fl_eq(eval_dotp_left[k] + eval_dotp_right[k], inst_evals[k]), VPIN_ERR_PROVER, "sparse evaluation mismatch");
  }
  std::vector<fl_t> rand_ops, rand_mem;
  phase("product_layer_setup", t0);
  double t1 = now_ms();
  BatchedS proof_ops = P.batched_prove(ops_ptr, N, dotp, ops_evals, &rand_ops, deal_ops ? kShardMinLayer : kRoundDealMin);
  phase("product_circuits_ops", t1);
  t1 = now_ms();
  BatchedS proof_mem = P.batched_prove(mem_ptr, M, {}, mem_evals, &rand_mem, deal_mem ? kShardMinLayer : kRoundDealMin);
  phase("product_circuits_mem", t1);
  t1 = now_ms();

  // HashLayerProof::prove (:740-849)
  t.protocol_name("Sparse polynomial hash layer proof");
  DevVec<fl_t> eq_ops = P.eq_table(rand_ops), eq_mem = P.eq_table(rand_mem);
  // `count` dot products <table_i, eq> over tables of `len` elements. Several GPUs: every rank takes one contiguous slice of the
  // index range, the partial sums (count elements per rank) are all-gathered and added on the host in rank order
  auto dots_vs = [&](const fl_t *base, size_t count, size_t len, const fl_t *eq) {
    const bool split = shard_sumcheck_enabled(ctx) && len >= ((size_t)1 << 18) && len % (size_t)world == 0;
    const size_t part = split ? len / world : len, o = split ? (size_t)rank * part : 0;
    std::vector<const fl_t *> hp(count);
    for (size_t i = 0; i < count; i++) hp[i] = base + i * len + o;
    DevVec<const fl_t *> dp(count, st);
    dp.upload(hp.data(), count);
    DevVec<fl_t> dout(count * (split ? world : 1), st);
    fl_t *mine = dout.p + (split ? (size_t)rank * count : 0);
    {
      ProfScope ps(ctx, PROF_DOT, (double)count * part, 32.0 * (count + 1) * part, 2);
      launch_dot_multi(dp.p, eq + o, (int)count, part, mine, ctx->d_partials.p, st);
    }
    if (split) dist_allgather_inplace(ctx, dout.p, count * sizeof(fl_t));
    std::vector<fl_t> all(count * (split ? world : 1)), out(count);
    dout.download(all.data(), all.size());
    ctx->sync();
    for (size_t i = 0; i < count; i++) {
      out[i] = all[i];
      for (int r = 1; split && r < world; r++) out[i] = out[i] + all[(size_t)r * count + i];
    }
    return out;
  };
  std::vector<fl_t> eval_derefs = dots_vs(derefs.p, 6, N, eq_ops.p);  // row A,B,C then col A,B,C
  std::vector<fl_t> eval_row_ops_val(eval_derefs.begin(), eval_derefs.begin() + 3), eval_col_ops_val(eval_derefs.begin() + 3, eval_derefs.end());
  // n-to-1 reduction shared by DerefsEvalProof::prove_single (:90-133) and the ops / mem openings (:780-838)
  auto reduce_n_to_1 = [&](std::vector<fl_t> evals, const char *chal_label, const std::vector<fl_t> &point, std::vector<fl_t> *r_joint) {
    std::vector<fl_t> challenges = t.challenge_vector(chal_label, math_log2(evals.size()));
    for (size_t i = challenges.size(); i-- > 0;) {  // bound_poly_var_bot in reverse challenge order
      size_t half = evals.size() / 2;
      for (size_t j = 0; j < half; j++) evals[j] = evals[2 * j] + challenges[i] * (evals[2 * j + 1] - evals[2 * j]);
      evals.resize(half);
    }
    *r_joint = challenges;
    r_joint->insert(r_joint->end(), point.begin(), point.end());
    return evals[0];
  };
  t.protocol_name("Derefs evaluation proof");
  DotLogS proof_derefs;
  {
    std::vector<fl_t> evals = eval_derefs;
    evals.resize(next_pow2(evals.size()), fl_zero());
    t.scalars("evals_ops_val", evals);
    std::vector<fl_t> r_joint;
    fl_t joint = reduce_n_to_1(evals, "challenge_combine_n_to_one", rand_ops, &r_joint);
    t.scalar("joint_claim_eval", joint);
    proof_derefs = P.polyeval_prove(dpc, *g.eval_label, derefs.p, nullptr, r_joint, joint, fl_zero(), nullptr);
  }
  std::vector<fl_t> evals15 = dots_vs(dec.comb_ops.p, 15, N, eq_ops.p);  // row addr, row read-ts, col addr, col read-ts, val (x3 each)
  std::vector<fl_t> eval_audit = dots_vs(dec.comb_mem.p, 2, M, eq_mem.p);
  DotLogS proof_ops_open, proof_mem_open;
  {
    std::vector<fl_t> evals = evals15;
    evals.resize(next_pow2(evals.size()), fl_zero());
    t.scalars("claim_evals_ops", evals);
    std::vector<fl_t> r_joint;
    fl_t joint = reduce_n_to_1(evals, "challenge_combine_n_to_one", rand_ops, &r_joint);
    t.scalar("joint_claim_eval_ops", joint);
    proof_ops_open = P.polyeval_prove(g.ops_pc, *g.eval_label, dec.comb_ops.p, nullptr, r_joint, joint, fl_zero(), nullptr);
  }
  {
    t.scalars("claim_evals_mem", eval_audit);
    std::vector<fl_t> r_joint;
    fl_t joint = reduce_n_to_1(eval_audit, "challenge_combine_two_to_one", rand_mem, &r_joint);
    t.scalar("joint_claim_eval_mem", joint);
    proof_mem_open = P.polyeval_prove(g.mem_pc, *g.eval_label, dec.comb_mem.p, nullptr, r_joint, joint, fl_zero(), nullptr);
  }
  phase("hash_layer_proof", t1);
  phase("evalproof_layered_network", t0);
  phase("R1CSEvalProof::prove", t_eval);
  ctx->phases.push_back({"batched_wait", P.t_b_wait});
  ctx->phases.push_back({"batched_small_wait", P.t_b_small_wait});
  ctx->phases.push_back({"batched_n_small", (double)P.n_b_small});
  ctx->phases.push_back({"batched_n_rounds", (double)P.n_b_rounds});
  ctx->phases.push_back({"batched_host", P.t_b_host});
  ctx->phases.push_back({"batched_launch", P.t_b_launch});
  ctx->phases.push_back({"bullet_pre(4 proofs)", P.t_bullet_pre});
  ctx->phases.push_back({"bullet_rounds_gpu(4 proofs)", P.t_bullet_gpu});
  ctx->phases.push_back({"bullet_rounds_host(4 proofs)", P.t_bullet_host});
  phase("SNARK::prove", t_prove);

  // ---------------- bincode(SNARK) (SP/lib.rs:330-338 and the nested types) ----------------
  Bin o;
  // R1CSProof (SP/r1csproof.rs:21-47)
  {
    std::vector<Comp> cv(spc.L);
    memcpy(cv.data(), wit.comm.data(), 32 * spc.L);
    o.comps(cv);
  }
  put(o, sc1);
  o.comp(comm_Az_claim); o.comp(comm_Bz_claim); o.comp(comm_Cz_claim); o.comp(comm_prod_Az_Bz_claims);
  put(o, pok_Cz_claim);
  put(o, proof_prod);
  put(o, proof_eq_sc_phase1);
  put(o, sc2);
  o.comp(comm_vars_at_ry);
  put(o, proof_eval_vars_at_ry);
  put(o, proof_eq_sc_phase2);
  // inst_evals
  for (int k = 0; k < 3; k++) o.fl(inst_evals[k]);
  // R1CSEvalProof { SparseMatPolyEvalProof { comm_derefs, poly_eval_network_proof { proof_prod_layer, proof_hash_layer } } }
  o.comps(comm_derefs);
  // ProductLayerProof (SP/sparse_mlpoly.rs:1035-1042)
  o.fl(er.init); o.fls(er.read); o.fls(er.write); o.fl(er.audit);
  o.fl(ec.init); o.fls(ec.read); o.fls(ec.write); o.fl(ec.audit);
  o.fls(eval_dotp_left); o.fls(eval_dotp_right);
  put(o, proof_mem);
  put(o, proof_ops);
  // HashLayerProof (SP/sparse_mlpoly.rs:698-707)
  std::vector<fl_t> row_addr(evals15.begin(), evals15.begin() + 3), row_rts(evals15.begin() + 3, evals15.begin() + 6);
  std::vector<fl_t> col_addr(evals15.begin() + 6, evals15.begin() + 9), col_rts(evals15.begin() + 9, evals15.begin() + 12);
  std::vector<fl_t> vals(evals15.begin() + 12, evals15.begin() + 15);
  o.fls(row_addr); o.fls(row_rts); o.fl(eval_audit[0]);
  o.fls(col_addr); o.fls(col_rts); o.fl(eval_audit[1]);
  o.fls(vals);
  o.fls(eval_row_ops_val); o.fls(eval_col_ops_val);
  put(o, proof_ops_open);
  put(o, proof_mem_open);
  put(o, proof_derefs);
  return std::move(o.b);
}

// ------------------------------------------------------------------------------------------------ integer roofline
// Peak rate of 32 x 32 + 64 -> 64 multiply-accumulates, measured live. Two forms, both checked in SASS (cuobjdump) to contain
// what they claim, with one factor changing every iteration so that nothing can be hoisted out of the loop:
//   form 0: IMAD.WIDE.U32 Rd, Ra, Rb, RZ  + LOP3 pair (product on the multiply pipe, combined into the accumulator on the ALU pipe)
//   form 1: IMAD.WIDE.U32 Rd, Ra, Rb, Rd  (single-instruction multiply-accumulate, the form the field arithmetic uses)
// The round-1 kernel that stood here multiplied two loop-invariant registers: ptxas hoisted the product and the loop timed
// IADD3 / IADD3.X pairs - 18.4 T "MAC"/s without a single multiply. IMAD.WIDE is a half-rate instruction on sm_100a: both
// forms top out near 32 lanes / clk / SM (8.0 and 7.6 T/s on a B200 at 1.965 GHz; scripts/ubench/imad_rates2.cu).
template <int kForm>
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t *out, int iters, uint32_t seed) {
  uint32_t x = threadIdx.x * 2654435761u + seed, y = x ^ 0x9e3779b9u;
  uint64_t v[12];
  uint32_t av[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { v[i] = x + i; av[i] = x * (2 * i + 3); }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
      if (kForm == 0) {
        uint64_t p;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(av[i]), "r"(y));
        v[i] ^= p;
      } else {
        uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(av[i]), "r"(y));
        v[i] = ((uint64_t)hi << 32) | lo;
      }
    }
    y += 0x9e3779b1u;
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) s ^= v[i];
  if (s == 0x123456789abcdefull) out[0] = (uint32_t)s;
}
void measure_imad_peaks(Ctx *ctx, double forms[2]) {
  cudaStream_t st = ctx->st;
  DevVec<uint32_t> out(1, st);
  int blocks = 148 * 8, iters = 4096;
  cudaEvent_t e0, e1;
  VPIN_CUDA(cudaEventCreate(&e0));
  VPIN_CUDA(cudaEventCreate(&e1));
  for (int form = 0; form < 2; form++) {
    auto launch = [&](int n, uint32_t seed) {
      ++g_kernel_launches;
      if (form == 0) k_imad_peak<0><<<blocks, 256, 0, st>>>(out.p, n, seed);
      else k_imad_peak<1><<<blocks, 256, 0, st>>>(out.p, n, seed);
    };
    launch(64, 1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
      VPIN_CUDA(cudaEventRecord(e0, st));
      launch(iters, rep);
      VPIN_CUDA(cudaEventRecord(e1, st));
      VPIN_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      VPIN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      double macs = (double)blocks * 256 * iters * 12 / (ms * 1e-3);
      if (macs > best) best = macs;
    }
    forms[form] = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}
double measure_imad_peak(Ctx *ctx) {  // the better of the two forms
  double f[2];
  measure_imad_peaks(ctx, f);
  return f[0] > f[1] ? f[0] : f[1];
}

}  // namespace vpin
