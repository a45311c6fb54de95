// Host orchestration of the B200 Spartan prover: public parameters, SNARK::encode, my_lib_prove.
// The Fiat-Shamir transcript and the O(1)-size sigma protocols run on the host; every table-sized operation is a kernel.
#pragma once
#include "core.cuh"

namespace vpin {

// PolyCommitmentGens::new(num_vars, label) (Spartan/src/dense_mlpoly.rs:36-40) -> DotProductProofGens::new(R, label)
// (Spartan/src/nizk/mod.rs:421-424): gens_n.G = stream[0..R), gens_1.G[0] = stream[R], h = stream[R+1].
struct PcGens {
  size_t ell = 0, L = 0, R = 0;
  size_t g1_index = 0, h_index = 0;
  const HostBase *g1 = nullptr, *h = nullptr;  // owned by the LabelGens of the stream
};

// SNARKGens (Spartan/src/lib.rs:295-327)
struct vpin_gens_impl {
  std::shared_ptr<LabelGens> sat_label, eval_label;  // streams "gens_r1cs_sat" / "gens_r1cs_eval" + fixed-base tables
  PcGens sat_pc;                                     // gens_r1cs_sat.gens_pc ; gens_sc.gens_1 is its gens_1
  const HostBase *sat_g[5];                               // stream[0..5): gens_3 = (G[0..3), h = G[3]); gens_4 = (G[0..4), h = G[4])
  PcGens ops_pc, mem_pc, derefs_pc;                  // gens_r1cs_eval.gens.{gens_ops, gens_mem, gens_derefs}
};
typedef vpin_gens_impl SnarkGens;

// MultiSparseMatPolynomialAsDense (Spartan/src/sparse_mlpoly.rs:285-292) resident in HBM
struct vpin_decomm_impl {
  size_t N = 0, M = 0;  // ops per matrix (padded nnz), memory cells
  size_t num_cons = 0, num_vars = 0, num_inputs = 0;  // of the instance (the commitment's bincode carries them)
  struct U32View { uint32_t *p = nullptr; };
  DevVec<uint32_t> row_addr_all, col_addr_all, row_read_ts_all, col_read_ts_all;  // 3N each: matrices A | B | C
  U32View row_addr[3], col_addr[3], row_read_ts[3], col_read_ts[3];                // views into the arrays above
  DevVec<uint32_t> row_audit_ts, col_audit_ts;
  DevVec<fl_t> comb_ops;  // 16 N: row-addr A,B,C | row-read-ts A,B,C | col-addr A,B,C | col-read-ts A,B,C | val A,B,C | 0
  DevVec<fl_t> comb_mem;  // 2 M: row audit-ts | col audit-ts
  const fl_t *val(int k) const { return comb_ops.p + (12 + k) * N; }
};
typedef vpin_decomm_impl Decomm;

struct vpin_witness_impl {
  size_t n_vars = 0;
  DevVec<fl_t> d_vars, d_blinds;
  std::vector<uint8_t> comm;  // L x 32 compressed rows (comm_vars)
};
typedef vpin_witness_impl Witness;

std::unique_ptr<SnarkGens> snark_gens_create(Ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs, uint64_t num_nz_entries);
std::unique_ptr<Decomm> snark_encode(Ctx *ctx, const Instance &inst, const SnarkGens &gens, std::vector<uint8_t> *comm_bytes);
// the two halves of SNARK::encode: the dense representation (timestamps by a radix sort, comb_ops, comb_mem: everything
// my_lib_prove reads - vPIN's prover never appends the computation commitment to its transcript, VP/commit_test.rs:75) and the two
// Hyrax commitments over it (the ComputationCommitment the verifier gets). The second half depends on nothing but the first and
// nothing in the proof depends on it, so a driver may run it on a background context while the proof is under way.
std::unique_ptr<Decomm> snark_encode_tables(Ctx *ctx, const Instance &inst, const SnarkGens &gens);
std::vector<uint8_t> snark_encode_commit(Ctx *ctx, const Decomm &decomm, const SnarkGens &gens);
std::vector<uint8_t> snark_prove(Ctx *ctx, const Instance &inst, const Decomm &decomm, const Witness &w, const std::vector<fl_t> &inputs,
                                 const SnarkGens &gens, const uint8_t *label, size_t label_len, const fl_t &tape_seed);
bool instance_is_sat(Ctx *ctx, const Instance &inst, const uint8_t *vars32, uint64_t n_vars, const uint8_t *inputs32, uint64_t n_inputs);

static inline CsrDev csr_of(const MatrixDev &m, size_t rows) {
  return CsrDev{m.csr_ptr.p, m.csr_col.p, m.csr_val.p, rows, m.csr_code.p, m.long_rows.p, m.n_long_rows};
}
static inline CscDev csc_of(const MatrixDev &m, size_t cols) {
  return CscDev{m.csc_ptr.p, m.csc_row.p, m.csc_val.p, cols, m.long_cols.p, m.n_long, m.csc_code.p};
}
static inline size_t eq_tmp_elems(size_t ell) {
  size_t a = (size_t)3 << ((ell + 1) / 2);
  return a < 4096 ? 4096 : a;
}
double measure_imad_peak(Ctx *ctx);
void measure_imad_peaks(Ctx *ctx, double forms[2]);

// packed product tree of n leaves stored at tree[0..n): 2n - 2 elements (prover.cu)
void build_tree(Ctx *ctx, fl_t *tree, size_t n, cudaStream_t st);
void build_trees(Ctx *ctx, const std::vector<fl_t *> &trees, size_t n, cudaStream_t st);

// builders.cu — vPIN_proof_generation/src/point_mult.rs, point_addition.rs
void point_mult_dims(uint64_t m, uint64_t dims_out[4]);
void point_add_dims(uint64_t n, uint64_t dims_out[4]);
std::unique_ptr<Instance> build_point_mult(Ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32, const uint8_t *py32,
                                           uint64_t dims_out[4], uint8_t *vars_para32, uint8_t *vars_input32, uint8_t *vars32,
                                           uint8_t *inputs32);
// device-resident variant: the three assignments are written in Montgomery form into caller-provided device arrays of
// `padded` elements each (zero beyond num_vars)
std::unique_ptr<Instance> build_point_mult_dev(Ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32, const uint8_t *py32,
                                               uint64_t dims_out[4], fl_t *d_para, fl_t *d_input, fl_t *d_vars, size_t padded,
                                               uint8_t *inputs32);
std::unique_ptr<Instance> build_point_add(Ctx *ctx, uint64_t n, const uint8_t *px32, const uint8_t *py32, const uint8_t *rx32,
                                          const uint8_t *ry32, const int64_t *rz_flags, uint64_t dims_out[4], uint8_t *vars_para32,
                                          uint8_t *vars_input32, uint8_t *vars32);

}  // namespace vpin
