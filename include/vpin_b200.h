/* vpin_b200.h — C ABI of the B200-native Spartan prover backend for vPIN's proof_generation crate.
 *
 * Drop-in boundary: the Rust functions vPIN's driver calls (vPIN_proof_generation/src/proof_point_add.rs:39-107,
 * proof_point_mult.rs:39-107) — SNARKGens::new, Instance::new, SNARK::encode, DensePolynomial::commit,
 * my_dense_mlpoly_commit, my_lib_prove — each get one entry point here with the same argument meaning.
 * A Rust shim binds these with `extern "C"`; in this repository the same symbols are bound by
 * vpin_b200/api.py through ctypes. "SP/" = Spartan/src/, "VP/" = vPIN_proof_generation/src/ in the reference tree.
 *
 * Conventions
 *   - scalars cross the ABI as 32-byte little-endian CANONICAL values (what Scalar::to_bytes / Instance::new use),
 *     never Montgomery form; group elements as 32-byte compressed ristretto255 (CompressedRistretto).
 *   - proofs and commitments-to-computations are bincode 1.3.3 byte strings identical to bincode::serialize(&SNARK) /
 *     (&ComputationCommitment) in the reference (VP/proof_point_add.rs:96).
 *   - the caller owns every input buffer for the duration of the call; outputs are written to caller buffers with an
 *     explicit capacity and a length out-parameter. Opaque handles own device (HBM) state.
 *   - every function returns a vpin_status; no function aborts. There is no CPU fallback: if no CUDA device is
 *     usable, vpin_ctx_create fails with VPIN_ERR_CUDA.
 *   - one context per host thread (the reference API is single-threaded: &mut Transcript, &mut RandomTape).
 *   - randomness: the reference seeds its RandomTapes from OsRng (SP/random.rs:16-18). Here the 32-byte `init_randomness`
 *     scalar of each tape is an argument: pass NULL and the library draws it from the operating system's CSPRNG (getrandom,
 *     64 bytes reduced mod l - what the reference does); pass 32 canonical bytes to reproduce a run (tests, parity vectors).
 *     A seed is ONE-TIME: the Hyrax blinds and every sigma-protocol nonce derive from it, so proving two different witnesses
 *     under the same seed leaks them. The fixed seeds in this repository's tests and bench exist for byte comparison only.
 */
#ifndef VPIN_B200_H
#define VPIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t vpin_status;
enum {
  VPIN_OK = 0,
  VPIN_ERR_INVALID_SCALAR = 1,    /* R1CSError::InvalidScalar            SP/errors.rs:42 */
  VPIN_ERR_INVALID_INDEX = 2,     /* R1CSError::InvalidIndex             SP/errors.rs:44 */
  VPIN_ERR_INVALID_NUM_INPUTS = 3,/* R1CSError::InvalidNumberOfInputs    SP/errors.rs:40 */
  VPIN_ERR_SIZE_MISMATCH = 4,     /* the reference's assert_eq! on lengths (e.g. SP/commitments.rs:95) */
  VPIN_ERR_BUFFER_TOO_SMALL = 5,
  VPIN_ERR_CUDA = 6,
  VPIN_ERR_OOM = 7,
  VPIN_ERR_PROVER = 8,            /* a prover-side assert! of the reference failed (unsatisfied witness, ...) */
  VPIN_ERR_BAD_ARGUMENT = 9,
  VPIN_ERR_IO = 10                /* a witness file could not be read / written / parsed (the reference's .expect panics, VP/load_data.rs:14-17) */
};

typedef struct vpin_ctx vpin_ctx;
typedef struct vpin_gens vpin_gens;         /* SNARKGens                    SP/lib.rs:295-327 */
typedef struct vpin_instance vpin_instance; /* Instance                     SP/lib.rs:130-244 */
typedef struct vpin_decomm vpin_decomm;     /* ComputationDecommitment      SP/lib.rs:66-69   */
typedef struct vpin_witness vpin_witness;   /* (vars, poly_vars, comm_vars, blinds_vars) resident in HBM */

/* one COO triple of Instance::new: (row, col, 32-byte LE canonical value)   SP/lib.rs:142-144 */
typedef struct vpin_coo_entry {
  uint64_t row;
  uint64_t col;
  uint8_t val[32];
} vpin_coo_entry;

/* ---- context ---------------------------------------------------------------------------------------------- */
vpin_status vpin_ctx_create(int32_t cuda_device, vpin_ctx **out);
/* high_priority > 0: the context's stream is created at the device's most urgent stream priority, so that when several
 * contexts share a GPU (a network's independent instances are proved concurrently, one context each) the thread blocks
 * of this one are scheduled first — give it to the instance on the critical path (the point-mult proof). */
vpin_status vpin_ctx_create_ex(int32_t cuda_device, int32_t high_priority, vpin_ctx **out);
/* high_priority < 0: a BACKGROUND context - the device's lowest stream priority and at most one resident MSM block per SM.
 * For work nothing waits for: vpin_encode_commit (below) while my_lib_prove runs on another context of the same device. */
void vpin_ctx_destroy(vpin_ctx *ctx);
/* Multi-GPU, one process (and one context) per GPU. Rank 0 calls vpin_nccl_unique_id and ships the 128 bytes to the other
 * ranks by any means (bench.py: a torch.distributed broadcast); every rank then calls vpin_ctx_init_distributed. After
 * that every rank runs the SAME calls with the SAME inputs (the Fiat-Shamir transcript is replayed identically on every
 * rank, so no challenge is ever broadcast); the L independent rows of every Hyrax commitment (SP/dense_mlpoly.rs:160-175,
 * the reference's one rayon par_iter) are split across ranks and exchanged with one in-place NCCL all-gather of 32 B
 * per row. vpin_shard_rows tells which rows a rank owns (whole range when the grid is too small to shard).
 * With VPIN_SHARD_SUMCHECK=1 in the environment the instances of the large batched product-circuit sumcheck layers are dealt to the
 * ranks as well (one NCCL all-gather of <= 3 KB per round); the proof bytes do not change. */
vpin_status vpin_nccl_unique_id(uint8_t id_out[128]);
vpin_status vpin_ctx_init_distributed(vpin_ctx *ctx, int32_t rank, int32_t world, const uint8_t nccl_id[128]);
void vpin_shard_rows(uint64_t rows, int32_t rank, int32_t world, uint64_t *r0, uint64_t *r1, int32_t *sharded);
/* sharded sumcheck rounds for the proofs of this context: 1 = on, 0 = off, -1 = what VPIN_SHARD_SUMCHECK says (the default).
 * Every rank of the communicator must choose the same value. */
vpin_status vpin_ctx_set_shard_sumcheck(vpin_ctx *ctx, int32_t on);
/* message of the last failure on this context (never NULL) */
const char *vpin_last_error(const vpin_ctx *ctx);
/* number of this library's kernels launched on the context so far (bench.py's gpu_launches) */
uint64_t vpin_kernel_launches(const vpin_ctx *ctx); /* kernels launched by calls on this context (ctx == NULL: by the process) */

/* ---- public parameters: SNARKGens::new(num_cons, num_vars, num_inputs, num_nz_entries)   SP/lib.rs:305 ---------- */
vpin_status vpin_gens_create(vpin_ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs,
                             uint64_t num_nz_entries, vpin_gens **out);
void vpin_gens_destroy(vpin_gens *gens);
/* L (rows) and R (columns) of the Hyrax grid of the witness polynomial (SP/dense_mlpoly.rs:96-98) */
vpin_status vpin_gens_witness_grid(const vpin_gens *gens, uint64_t *L, uint64_t *R);

/* ---- Instance::new(num_cons, num_vars, num_inputs, &A, &B, &C)   SP/lib.rs:138-244 ------------------------------ */
vpin_status vpin_instance_create(vpin_ctx *ctx, uint64_t num_cons, uint64_t num_vars, uint64_t num_inputs,
                                 const vpin_coo_entry *A, uint64_t nA, const vpin_coo_entry *B, uint64_t nB,
                                 const vpin_coo_entry *C, uint64_t nC, vpin_instance **out);
void vpin_instance_destroy(vpin_instance *inst);
/* padded sizes (inst.inst.get_num_cons / get_num_vars) */
vpin_status vpin_instance_dims(const vpin_instance *inst, uint64_t *num_cons_padded, uint64_t *num_vars_padded,
                               uint64_t *num_inputs);
/* Instance::is_sat(&vars, &inputs)   SP/lib.rs:247-276 ; vars may be shorter than the padded size */
vpin_status vpin_instance_is_sat(vpin_ctx *ctx, const vpin_instance *inst, const uint8_t *vars32, uint64_t n_vars,
                                 const uint8_t *inputs32, uint64_t n_inputs, int32_t *sat);

/* the instance's COO triples back in Instance::new's format (unpadded column indices, canonical values) */
vpin_status vpin_instance_nnz(const vpin_instance *inst, uint64_t nnz_out[3]);
vpin_status vpin_instance_export_coo(vpin_ctx *ctx, const vpin_instance *inst, uint64_t num_vars_unpadded, int32_t which,
                                     vpin_coo_entry *out);

/* ---- SNARK::encode(&inst, &gens)   SP/lib.rs:347-358 -------------------------------------------------------------
 * comm_out receives bincode(ComputationCommitment); *decomm keeps the dense representation in HBM. */
vpin_status vpin_encode(vpin_ctx *ctx, const vpin_instance *inst, const vpin_gens *gens, uint8_t *comm_out,
                        uint64_t comm_cap, uint64_t *comm_len, vpin_decomm **decomm);
/* The two halves of SNARK::encode, for a driver that overlaps them with the rest of its flow:
 *  vpin_encode_tables: the dense representation my_lib_prove reads (SP/sparse_mlpoly.rs:382-438 MultiSparseMatPolynomialAsDense,
 *    AddrTimestamps::new :232-265) - no commitment. The handle may be used by a proof on any context of the device.
 *  vpin_encode_commit: the two Hyrax commitments over it (SP/sparse_mlpoly.rs:500-520) -> bincode(ComputationCommitment).
 * vPIN's my_lib_prove never appends the computation commitment to its transcript (VP/commit_test.rs:75, unlike SP/lib.rs:377), so
 * the proof does not wait for the second half: run it on a background context (vpin_ctx_create_ex(.., -1, ..)) under the proof.
 * vpin_encode == vpin_encode_tables + vpin_encode_commit on one context. */
vpin_status vpin_encode_tables(vpin_ctx *ctx, const vpin_instance *inst, const vpin_gens *gens, vpin_decomm **decomm);
vpin_status vpin_encode_commit(vpin_ctx *ctx, const vpin_decomm *decomm, const vpin_gens *gens, uint8_t *comm_out,
                               uint64_t comm_cap, uint64_t *comm_len);
void vpin_decomm_destroy(vpin_decomm *decomm);

/* ---- DensePolynomial::commit(&gens.gens_r1cs_sat.gens_pc, Some(&mut tape))   SP/dense_mlpoly.rs:193-218 ----------
 * Z: n = 2^ell canonical scalars (the padded assignment). Blinds are drawn from a RandomTape exactly as the
 * reference does: `tape_state` is an opaque 256-byte buffer initialised by vpin_tape_init and advanced by each call,
 * so two consecutive commits on one tape reproduce VP/proof_point_add.rs:44-52. Pass tape_state = NULL for zero
 * blinds. points_out: L x 32 bytes, blinds_out: L x 32 bytes (may be NULL).
 * init_randomness32 == NULL: fresh randomness from the OS (production use); otherwise 32 canonical bytes (one-time!). */
vpin_status vpin_tape_init(uint8_t tape_state[256], const uint8_t *name, uint64_t name_len,
                           const uint8_t init_randomness32[32]);
vpin_status vpin_poly_commit(vpin_ctx *ctx, const vpin_gens *gens, const uint8_t *Z32, uint64_t n,
                             uint8_t *tape_state, uint8_t *points_out, uint8_t *blinds_out);
/* my_dense_mlpoly_commit(&poly, &gens_pc, blind_1, blind_2)   VP/commit_test.rs:27-57 (blinds = blind_1 + blind_2) */
vpin_status vpin_poly_commit_with_blinds(vpin_ctx *ctx, const vpin_gens *gens, const uint8_t *Z32, uint64_t n,
                                         const uint8_t *blind1_32, const uint8_t *blind2_32, uint64_t L,
                                         uint8_t *points_out, uint8_t *blinds_out);
/* row-wise C1[i] + C2[i] on compressed points (VP/proof_point_add.rs:75-80, VP/commit_test.rs:355-358).
 * Fails with VPIN_ERR_BAD_ARGUMENT if a point does not decompress. */
vpin_status vpin_commitments_add(vpin_ctx *ctx, const uint8_t *c1, const uint8_t *c2, uint64_t L, uint8_t *out);

/* ---- my_lib_prove(&inst, &decomm, vars, &inputs, &gens, &mut transcript, poly_vars, comm_vars, blinds_vars)
 *      VP/commit_test.rs:59-133 ------------------------------------------------------------------------------------
 * vars32: the padded assignment (num_vars_padded scalars) — also the evaluations of poly_vars.
 * transcript_label: the label of Transcript::new (b"snark_example" in VP/proof_point_add.rs:83).
 * tape_seed32: init_randomness of RandomTape::new(b"proof") (VP/commit_test.rs:74); NULL = drawn from the OS (see Conventions).
 * proof_out receives bincode(SNARK). */
vpin_status vpin_prove(vpin_ctx *ctx, const vpin_instance *inst, const vpin_decomm *decomm, const uint8_t *vars32,
                       uint64_t n_vars, const uint8_t *inputs32, uint64_t n_inputs, const vpin_gens *gens,
                       const uint8_t *transcript_label, uint64_t label_len, const uint8_t *comm_vars_points,
                       const uint8_t *blinds_vars32, uint64_t L, const uint8_t tape_seed32[32], uint8_t *proof_out,
                       uint64_t proof_cap, uint64_t *proof_len);
/* same proof with the witness already resident in HBM (what bench.py times as `value`) */
vpin_status vpin_witness_upload(vpin_ctx *ctx, const vpin_gens *gens, const uint8_t *vars32, uint64_t n_vars,
                                const uint8_t *comm_vars_points, const uint8_t *blinds_vars32, uint64_t L,
                                vpin_witness **out);
void vpin_witness_destroy(vpin_witness *w);
vpin_status vpin_prove_resident(vpin_ctx *ctx, const vpin_instance *inst, const vpin_decomm *decomm,
                                const vpin_witness *w, const uint8_t *inputs32, uint64_t n_inputs,
                                const vpin_gens *gens, const uint8_t *transcript_label, uint64_t label_len,
                                const uint8_t tape_seed32[32], uint8_t *proof_out, uint64_t proof_cap,
                                uint64_t *proof_len);
/* milliseconds (CUDA events + host clock) of the last vpin_prove* under the reference's timer labels
 * (SP/timer.rs; VP/commit_test.rs:70,101,147,153,159,249,283; SP/sparse_mlpoly.rs:1491,1504,1514).
 * names_out[i] points at static strings; returns the number of phases written (<= cap). */
uint32_t vpin_last_phase_times(const vpin_ctx *ctx, const char **names_out, double *ms_out, uint32_t cap);

/* ---- vPIN's R1CS builders (fixture generators; VP/point_addition.rs:5-326, VP/point_mult.rs:7-704) ---------------
 * Inputs are what VP/load_data.rs / load_data_add.rs read from the JSON files. Outputs: the instance plus the three
 * UNPADDED assignments (vars_para, vars_input, vars: num_vars x 32 bytes each), the inputs assignment, and the
 * (num_cons, num_vars, num_inputs, num_non_zero_entries) tuple the driver passes to SNARKGens::new. */
vpin_status vpin_build_point_mult(vpin_ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32,
                                  const uint8_t *py32, vpin_instance **inst, uint64_t dims_out[4],
                                  uint8_t *vars_para32, uint8_t *vars_input32, uint8_t *vars32, uint8_t *inputs32);
/* The same builder with the three assignments left in HBM: d_vars_para / d_vars_input / d_vars are device arrays of
 * `padded` 32-byte elements each (padded >= num_vars; normally the power of two the prover pads to, see
 * vpin_instance_dims), written in Montgomery form and zero beyond num_vars — exactly what vpin_dev_poly_commit* and
 * vpin_witness_from_device take. Witness expansion (VP/point_mult.rs:328-602, 256 field inversions per multiplication)
 * and COO emission (:85-322) both run on the device; the reference does them on one core inside its timed region
 * (VP/proof_point_mult.rs:24-101). */
vpin_status vpin_build_point_mult_device(vpin_ctx *ctx, uint64_t m, const uint64_t *weights_lo_hi, const uint8_t *px32,
                                         const uint8_t *py32, vpin_instance **inst, uint64_t dims_out[4],
                                         void *d_vars_para, void *d_vars_input, void *d_vars, uint64_t padded,
                                         uint8_t *inputs32);
vpin_status vpin_build_point_add(vpin_ctx *ctx, uint64_t n, const uint8_t *px32, const uint8_t *py32,
                                 const uint8_t *rx32, const uint8_t *ry32, const int64_t *rz_flags,
                                 vpin_instance **inst, uint64_t dims_out[4], uint8_t *vars_para32,
                                 uint8_t *vars_input32, uint8_t *vars32);
/* sizes only (to allocate the buffers above): dims_out = num_cons, num_vars, num_inputs, num_non_zero_entries */
void vpin_point_mult_dims(uint64_t m, uint64_t dims_out[4]);
void vpin_point_add_dims(uint64_t n, uint64_t dims_out[4]);

/* ---- kernel-level entry points (parity tests and roofline benches) -------------------------
 * Host-buffer forms: canonical scalars in, canonical scalars / compressed points out. */
/* sum_i s_i * G_i over the first n generators of the SHAKE256 stream `label` (SP/commitments.rs:20-38,
 * SP/group.rs:103-121) */
vpin_status vpin_msm(vpin_ctx *ctx, const char *label, const uint8_t *scalars32, uint64_t n, uint8_t out_point[32]);
/* Hyrax commitment of 2^ell scalars with PolyCommitmentGens::new(ell, label) (SP/dense_mlpoly.rs:36-40,160-175);
 * blinds32 may be NULL */
vpin_status vpin_hyrax_commit(vpin_ctx *ctx, const char *label, const uint8_t *Z32, uint64_t n, const uint8_t *blinds32,
                              uint8_t *points_out);
/* n+1 generators of MultiCommitGens::new(n, label), compressed (the last one is h) */
vpin_status vpin_derive_gens(vpin_ctx *ctx, const char *label, uint64_t n, uint8_t *points_out);
/* (Az, Bz, Cz) = inst.multiply_vec(z)   SP/r1csinstance.rs:272-286; z has 2*num_vars_padded entries */
vpin_status vpin_spmv_abc(vpin_ctx *ctx, const vpin_instance *inst, const uint8_t *z32, uint8_t *Az32, uint8_t *Bz32,
                          uint8_t *Cz32);
/* (A^T x, B^T x, C^T x) = inst.compute_eval_table_sparse(x)   SP/r1csinstance.rs:288-302 */
vpin_status vpin_spmv_t_abc(vpin_ctx *ctx, const vpin_instance *inst, const uint8_t *x32, uint8_t *At32, uint8_t *Bt32,
                            uint8_t *Ct32);
/* EqPolynomial::new(r).evals()   SP/dense_mlpoly.rs:78-94 */
vpin_status vpin_eq_evals(vpin_ctx *ctx, const uint8_t *r32, uint32_t ell, uint8_t *out32);
/* eval_point_0, _2, _3 of one round with comb = A*(B*C - D)   SP/sumcheck.rs:619-652, SP/r1csproof.rs:104-108 */
vpin_status vpin_sumcheck_cubic_round(vpin_ctx *ctx, const uint8_t *A32, const uint8_t *B32, const uint8_t *C32,
                                      const uint8_t *D32, uint64_t len, uint8_t out96[96]);
/* eval_point_0, _2 with comb = A*B   SP/sumcheck.rs:456-469 */
vpin_status vpin_sumcheck_quad_round(vpin_ctx *ctx, const uint8_t *A32, const uint8_t *B32, uint64_t len,
                                     uint8_t out64[64]);
/* eval_point_0, _2, _3 with comb = A*B*C   SP/sumcheck.rs:296-320 */
vpin_status vpin_sumcheck_cubic3_round(vpin_ctx *ctx, const uint8_t *A32, const uint8_t *B32, const uint8_t *C32,
                                       uint64_t len, uint8_t out96[96]);
/* A whole sumcheck driven with caller-supplied challenges through the prover's FUSED round kernels (bind with r_{j-1} and
 * evaluate round j in one launch, results in host-mapped memory): the device side of prove_cubic_with_additive_term
 * (degree 3: comb = A*(B*C - D), SP/sumcheck.rs:619-676) or prove_quad (degree 2: comb = A*B, :456-486; C32, D32 NULL).
 * len = 2^rounds scalars per table, r32 = rounds challenges. evals_out: rounds x degree scalars (eval_point_0, _2[, _3]
 * of every round); finals_out: the tables bound at all challenges (4 or 2 scalars = poly[0] after the last bind). */
vpin_status vpin_sumcheck_fused(vpin_ctx *ctx, uint32_t degree, const uint8_t *A32, const uint8_t *B32, const uint8_t *C32,
                                const uint8_t *D32, uint64_t len, const uint8_t *r32, uint8_t *evals_out, uint8_t *finals_out);
/* AddrTimestamps::new (SP/sparse_mlpoly.rs:232-265) for one side (rows or columns) of the three matrices: addr_k = n_k
 * addresses < M in COO order, padded with address 0 to N operations each (SP/sparse_mlpoly.rs:370-378). Outputs: the 3N
 * padded addresses and read timestamps (A | B | C) and the M audit timestamps, all u32. */
vpin_status vpin_spark_timestamps(vpin_ctx *ctx, const uint32_t *addr_a, uint64_t n_a, const uint32_t *addr_b, uint64_t n_b,
                                  const uint32_t *addr_c, uint64_t n_c, uint64_t N, uint64_t M, uint32_t *addr_out,
                                  uint32_t *read_ts_out, uint32_t *audit_ts_out);
/* bound_poly_var_top   SP/dense_mlpoly.rs:229-236; Z32 (len scalars) is overwritten, first len/2 are the result */
vpin_status vpin_bind_top(vpin_ctx *ctx, uint8_t *Z32, uint64_t len, const uint8_t r32[32]);
/* DensePolynomial::bound(L)   SP/dense_mlpoly.rs:220-227: out has R = 2^ceil(ell/2) scalars */
vpin_status vpin_bound(vpin_ctx *ctx, const uint8_t *Z32, uint64_t len, const uint8_t *L32, uint8_t *out32);
/* ProductCircuit::new (SP/product_tree.rs:36-56, compute_layer :18-34): the packed tree of n = 2^k leaves,
 * tree = [V_0 | V_1 | ...] with V_0 = leaves and V_{j+1}[i] = V_j[i] * V_j[i + |V_j|/2], down to the layer of length 2 (the two
 * factors of the root): 2n - 2 elements out. n >= 2. */
vpin_status vpin_product_tree(vpin_ctx *ctx, const uint8_t *leaves32, uint64_t n, uint8_t *tree32);
/* Layers::build_hash_layer (SP/sparse_mlpoly.rs:547-622) for n operations: read[i] = ts[i] gamma^2 + val[i] gamma + addr[i] - tau,
 * write[i] = read[i] + gamma^2 (write_ts = read_ts + 1). With addr[i] = i and ts = 0 / audit_ts it is the init / audit layer. */
vpin_status vpin_hash_layer(vpin_ctx *ctx, const uint32_t *addr, const uint8_t *val32, const uint32_t *ts, uint64_t n,
                            const uint8_t gamma32[32], const uint8_t tau32[32], uint8_t *read32, uint8_t *write32);
/* deref_mem (SP/sparse_mlpoly.rs:267-276): out[i] = mem[addr[i]] */
vpin_status vpin_deref_gather(vpin_ctx *ctx, const uint32_t *addr, uint64_t n, const uint8_t *mem32, uint64_t num_cells,
                              uint8_t *out32);

/* Device-resident forms for benchmarks: pointers are CUDA device pointers to 32-byte MONTGOMERY elements (the
 * in-HBM table format, identical to the reference's serde of Scalar, SP/scalar/ristretto255.rs:199-200).
 * They enqueue on the context stream; call vpin_sync (or record CUDA events on vpin_stream) to wait. */
void *vpin_stream(vpin_ctx *ctx); /* cudaStream_t */
vpin_status vpin_sync(vpin_ctx *ctx);
vpin_status vpin_dev_to_mont(vpin_ctx *ctx, const void *d_in, uint64_t n, void *d_out);
vpin_status vpin_dev_from_mont(vpin_ctx *ctx, const void *d_in, uint64_t n, void *d_out);
vpin_status vpin_dev_hyrax_commit(vpin_ctx *ctx, const char *label, const void *d_Z, uint64_t n, const void *d_blinds,
                                  void *d_points_out);
vpin_status vpin_dev_cubic_round(vpin_ctx *ctx, const void *dA, const void *dB, const void *dC, const void *dD,
                                 uint64_t len, void *d_out3);
vpin_status vpin_dev_quad_round(vpin_ctx *ctx, const void *dA, const void *dB, uint64_t len, void *d_out2);
vpin_status vpin_dev_bind_top(vpin_ctx *ctx, void *dZ, uint64_t len, const void *d_r);
vpin_status vpin_dev_eq_evals(vpin_ctx *ctx, const void *d_r, uint32_t ell, void *d_out);
vpin_status vpin_dev_spmv_abc(vpin_ctx *ctx, const vpin_instance *inst, const void *d_z, void *dAz, void *dBz, void *dCz);
/* HBM-resident forms of the three witness commitments of VP/proof_point_add.rs:44-80 (what bench.py times as `value`):
 * d_Z: n Montgomery scalars; outputs stay on the device (L x 32 B compressed points, L Montgomery blinds). */
vpin_status vpin_dev_poly_commit(vpin_ctx *ctx, const vpin_gens *gens, const void *d_Z, uint64_t n, uint8_t *tape_state,
                                 void *d_points_out, void *d_blinds_out);
vpin_status vpin_dev_poly_commit_with_blinds(vpin_ctx *ctx, const vpin_gens *gens, const void *d_Z, uint64_t n, const void *d_blind1,
                                             const void *d_blind2, void *d_points_out, void *d_blinds_out);
vpin_status vpin_dev_commitments_add(vpin_ctx *ctx, const void *d_c1, const void *d_c2, uint64_t L, void *d_out);
vpin_status vpin_witness_from_device(vpin_ctx *ctx, const vpin_gens *gens, const void *d_vars, uint64_t n_vars,
                                     const void *d_comm_points, const void *d_blinds, uint64_t L, vpin_witness **out);
/* Per-kernel-class device timing with CUDA events on the context stream (bench.py's roofline lines). Scopes that
 * process fewer than min_units elements are not timed. vpin_profile_read drains the events and returns the number of
 * classes written; *msm_madds_out = mixed additions executed by the MSM kernels since vpin_profile_enable. */
vpin_status vpin_profile_enable(vpin_ctx *ctx, int32_t on, double min_units);
uint32_t vpin_profile_read(vpin_ctx *ctx, const char **names_out, double *ms_out, uint64_t *launches_out, double *units_out,
                           double *bytes_out, uint32_t cap, uint64_t *msm_madds_out);
/* The integer roofline, measured live: 32 x 32 + 64 -> 64 multiply-accumulates per second of a dependency-free stream of real
 * IMAD.WIDE.U32 instructions (one factor changes every iteration, so nothing is hoisted). forms[0]: plain product (no addend)
 * combined into the accumulator on the ALU pipe; forms[1]: single-instruction multiply-accumulate (64-bit addend), the form
 * the field arithmetic uses. vpin_imad_peak returns the better of the two. IMAD.WIDE is a half-rate instruction on sm_100a:
 * ~8.0 / 7.6 T/s on a B200 (the round-1 figure of 18.4 T/s timed IADD3 pairs - its product had been hoisted by ptxas). */
vpin_status vpin_imad_peak(vpin_ctx *ctx, double *macs_per_second);
vpin_status vpin_imad_peak_forms(vpin_ctx *ctx, double forms[2]);

/* ---- native witness I/O (host only)   VP/load_data.rs:5-62, VP/load_data_add.rs:5-102 --------------------------------------
 * Reads what the Python side writes for the Rust driver: <root>/rust_files/<tag>/pointMult/{weight.json, point_mult_px_byte.json,
 * point_mult_py_byte.json} and <root>/rust_files/<tag>/pointAdd/point_add_{px,py,rx,ry,rz}_byte.json (the reference opens these
 * paths relative to its working directory, i.e. root = "."). weights: u128 as (low, high) u64 pairs; coordinates: 32 little-endian
 * bytes per row; rz[i] = 1 marks R_i = infinity. Call once with all output buffers NULL to learn the count, then with buffers of
 * `cap` entries. A binary sidecar <dir>/witness.bin (vpin_witness_json_to_bin) holding the same content is preferred when present:
 * LeNet layer 5's JSON is 3 x 6000 lists of 32 decimal integers. vpin_witness_last_error(): message of this thread's last failure. */
vpin_status vpin_load_point_mult(const char *root, const char *tag, uint64_t cap, uint64_t *count_out, uint64_t *weights_lo_hi, uint8_t *px32,
                                 uint8_t *py32);
vpin_status vpin_load_point_add(const char *root, const char *tag, uint64_t cap, uint64_t *count_out, uint8_t *px32, uint8_t *py32, uint8_t *rx32,
                                uint8_t *ry32, int64_t *rz);
/* writes witness.bin next to the JSON files of whichever of the two directories exist; *written_out: bit 0 pointMult, bit 1 pointAdd */
vpin_status vpin_witness_json_to_bin(const char *root, const char *tag, uint32_t *written_out);
const char *vpin_witness_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VPIN_B200_H */
