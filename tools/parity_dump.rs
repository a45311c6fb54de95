// parity_dump.rs — prints, from the REAL reference crates, the bytes this repository's CUDA prover and CPU oracle claim to
// reproduce: the computation commitment, the three witness commitments and the bincode proof of vPIN's driver flow
// (vPIN_proof_generation/src/proof_point_add.rs:39-98 == proof_point_mult.rs:39-98) under FIXED tape seeds.
//
// Why it exists: no reference test pins a commitment, a challenge or a proof byte, and the Rust
// toolchain is absent from the build image, so the composition (Merlin framing, label order, bincode field order) is pinned
// only by restatement. Wherever `cargo` and the crates are available this program closes that gap.
//
// How to run it (from the reference checkout, `REF` = src/proof_generation):
//   1. Determinism hook — the reference seeds every RandomTape from OsRng (Spartan/src/random.rs:14-21). Apply the five-line
//      patch below to REF/Spartan/src/random.rs (it changes nothing unless the environment variable is set):
//
//        pub fn new(name: &'static [u8]) -> Self {
//          let tape = {
//            let mut csprng: OsRng = OsRng;
//            let mut tape = Transcript::new(name);
//      +     let seed = std::env::var(format!("VPIN_TAPE_SEED_{}", name.iter().map(|b| format!("{:02x}", b)).collect::<String>()))
//      +       .ok().map(|h| { let mut b = [0u8; 32];
//      +         for i in 0..32 { b[i] = u8::from_str_radix(&h[2 * i..2 * i + 2], 16).unwrap(); }
//      +         Scalar::from_bytes(&b).unwrap() });
//      -     tape.append_scalar(b"init_randomness", &Scalar::random(&mut csprng));
//      +     tape.append_scalar(b"init_randomness", &seed.unwrap_or_else(|| Scalar::random(&mut csprng)));
//            tape
//          };
//
//   2. cp tools/parity_dump.rs REF/vPIN_proof_generation/src/bin/parity_dump.rs
//   3. Write the witness files of a named shape with this repository's generator (same seeds as tests/golden/make_golden.py):
//        python -c "from vpin_b200 import workloads as W; m, n = W.SHAPES['conv3']; \
//                   W.write_rust_files('REF/vPIN_proof_generation', 'conv3', mult=W.synth_point_mult(m), add=W.synth_point_add(n))"
//   4. cd REF/vPIN_proof_generation && \
//      VPIN_TAPE_SEED_02=<tape_seeds[0] of tests/golden/golden_named.json> \
//      VPIN_TAPE_SEED_70726f6f66=<tape_seeds[1]> \
//      cargo run --release --bin parity_dump -- conv3 out_dir        ("02" = the tape named [2u8], "70726f6f66" = b"proof")
//   5. python tools/parity_check.py out_dir conv3        (compares sha256 / lengths with tests/golden/golden_named.json)
//
// It writes <out_dir>/<kind>_{comm,comm_vars_para,comm_vars_input,comm_vars,proof}.bin for kind in {point_add, point_mult}.
#![allow(non_snake_case)]
extern crate curve25519_dalek;
extern crate libspartan;
extern crate merlin;

#[path = "../commit_test.rs"]
pub mod commit_test;
#[path = "../load_data.rs"]
pub mod load_data;
#[path = "../load_data_add.rs"]
pub mod load_data_add;
#[path = "../point_addition.rs"]
pub mod point_addition;
#[path = "../point_mult.rs"]
pub mod point_mult;

use commit_test::{my_dense_mlpoly_commit, my_lib_prove, my_lib_verify};
use libspartan::dense_mlpoly::{DensePolynomial, PolyCommitment};
use libspartan::random::RandomTape;
use libspartan::{ComputationCommitment, InputsAssignment, Instance, SNARKGens, VarsAssignment, SNARK};
use merlin::Transcript;
use std::fs;
use std::path::Path;

fn dump(dir: &Path, kind: &str, what: &str, bytes: &[u8]) {
  fs::write(dir.join(format!("{}_{}.bin", kind, what)), bytes).expect("cannot write output file");
  println!("{} {}: {} bytes", kind, what, bytes.len());
}

// PolyCommitment { C: Vec<CompressedRistretto> } as the L x 32 raw bytes the C ABI returns (no length prefix)
fn rows(c: &PolyCommitment) -> Vec<u8> {
  c.C.iter().flat_map(|p| p.as_bytes().to_vec()).collect()
}

#[allow(clippy::too_many_arguments)]
fn flow(
  dir: &Path,
  kind: &str,
  num_cons: usize,
  num_vars: usize,
  num_inputs: usize,
  num_non_zero_entries: usize,
  inst: Instance,
  padded_vars_para: VarsAssignment,
  padded_vars_input: VarsAssignment,
  padded_vars: VarsAssignment,
  assignment_inputs: InputsAssignment,
) {
  // the driver sequence, statement by statement (proof_point_add.rs:39-98)
  let gens = SNARKGens::new(num_cons, num_vars, num_inputs, num_non_zero_entries);
  let (comm, decomm): (ComputationCommitment, _) = SNARK::encode(&inst, &gens);
  let mut random_tape_1 = RandomTape::new(&[2u8]);
  let poly_vars_para = DensePolynomial::new(padded_vars_para.assignment.clone());
  let (comm_vars_para, blind_vars_para) = poly_vars_para.commit(&gens.gens_r1cs_sat.gens_pc, Some(&mut random_tape_1));
  let poly_vars_inputs = DensePolynomial::new(padded_vars_input.assignment.clone());
  let (comm_vars_input, blind_vars_input) = poly_vars_inputs.commit(&gens.gens_r1cs_sat.gens_pc, Some(&mut random_tape_1));
  let poly_vars = DensePolynomial::new(padded_vars.assignment.clone());
  let (comm_vars, blind_vars) =
    my_dense_mlpoly_commit(&poly_vars, &gens.gens_r1cs_sat.gens_pc, blind_vars_para.blinds, blind_vars_input.blinds);
  let mut combine_comm_vars = vec![];
  for i in 0..comm_vars_para.C.len() {
    combine_comm_vars.push((comm_vars_para.C[i].decompress().unwrap() + comm_vars_input.C[i].decompress().unwrap()).compress());
  }
  let combine_commitment = PolyCommitment { C: combine_comm_vars };

  dump(dir, kind, "comm", &bincode::serialize(&comm).expect("serialize comm"));
  dump(dir, kind, "comm_vars_para", &rows(&comm_vars_para));
  dump(dir, kind, "comm_vars_input", &rows(&comm_vars_input));
  dump(dir, kind, "comm_vars", &rows(&comm_vars));

  let mut prover_transcript = Transcript::new(b"snark_example");
  let proof = my_lib_prove(
    &inst,
    &decomm,
    padded_vars,
    &assignment_inputs,
    &gens,
    &mut prover_transcript,
    poly_vars,
    combine_commitment,
    blind_vars,
  );
  dump(dir, kind, "proof", &bincode::serialize(&proof).expect("serialize proof"));

  let mut verifier_transcript = Transcript::new(b"snark_example");
  assert!(my_lib_verify(proof, &comm, &assignment_inputs, &mut verifier_transcript, &gens, comm_vars_para, comm_vars_input).is_ok());
  println!("{}: verified", kind);
}

fn main() {
  let args: Vec<String> = std::env::args().collect();
  let network = if args.len() > 1 { args[1].as_str() } else { "conv3" };
  let out = if args.len() > 2 { args[2].clone() } else { String::from("parity_out") };
  fs::create_dir_all(&out).expect("cannot create the output directory");
  let dir = Path::new(&out);
  for name in ["02", "70726f6f66"] {
    if std::env::var(format!("VPIN_TAPE_SEED_{}", name)).is_err() {
      eprintln!("warning: VPIN_TAPE_SEED_{} is not set - the tapes draw from OsRng and the bytes will not be reproducible", name);
    }
  }
  {
    let (num_cons, num_vars, num_inputs, nnz, inst, vp, vi, v, inputs) = point_addition::point_addition(network);
    flow(dir, "point_add", num_cons, num_vars, num_inputs, nnz, inst, vp, vi, v, inputs);
  }
  if network != "L2" && network != "L4" {
    let (num_cons, num_vars, num_inputs, nnz, inst, vp, vi, v, inputs) = point_mult::point_mult(network);
    flow(dir, "point_mult", num_cons, num_vars, num_inputs, nnz, inst, vp, vi, v, inputs);
  }
}
