"""CPU: the oracle's prover/verifier on the reference's own round-trip shapes (Spartan/src/lib.rs:615-774,
r1csproof.rs:515-619), the committed golden digests, and the error paths. No GPU."""
import hashlib
import importlib.util
import json
import os

import pytest

import helpers as H
import oracle_lib as O
from vpin_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_flows.json")))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)


def test_tape_seeds_are_the_committed_ones():
    assert [s.hex() for s in W.tape_seeds()] == GOLD["tape_seeds"]


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"{c['kind']}-{'-'.join(str(v) for v in c['args'].values())}")
def test_oracle_reproduces_golden_digests(case):
    got = MG.run(case["kind"], case["args"])
    assert got == case


def test_prove_verify_round_trip_and_rejections():
    A, B, Cm, vp, vi, v, inputs = H.synthetic_r1cs(64, 32, 5, seed=64)
    built = O.build_custom(64, 32, 5, 64, A, B, Cm, vp, vi, v, inputs)
    sq, sp = W.tape_seeds()
    f = O.Flow(built, sq, sp, verify=True)
    assert f.verified
    dims = built.dims
    assert O.verify(dims, f.proof, f.comm, inputs, f.comm_vars_para, f.comm_vars_input) == 1
    # wrong public input, corrupted proof, swapped commitments: all rejected (my_lib_verify returns Err)
    bad_inputs = bytearray(inputs)
    bad_inputs[0] ^= 1
    assert O.verify(dims, f.proof, f.comm, bytes(bad_inputs), f.comm_vars_para, f.comm_vars_input) != 1
    for off in (40, len(f.proof) // 3, len(f.proof) - 40):
        bad = bytearray(f.proof)
        bad[off] ^= 0x10
        assert O.verify(dims, bytes(bad), f.comm, inputs, f.comm_vars_para, f.comm_vars_input) != 1
    assert O.verify(dims, f.proof, f.comm, inputs, f.comm_vars_input, f.comm_vars_input) != 1
    # different tape seeds change the proof (blinds) but not the computation commitment
    f2 = O.Flow(built, sp, sq, verify=True)
    assert f2.verified and f2.proof != f.proof and f2.comm == f.comm


def test_proof_is_deterministic_and_thread_count_independent():
    built = O.build_point_add(*W.synth_point_add(8, infinity_every=5))
    sq, sp = W.tape_seeds()
    a = O.Flow(built, sq, sp, verify=False, threads=1)
    b = O.Flow(built, sq, sp, verify=False, threads=4)
    assert a.proof == b.proof and a.comm == b.comm and a.comm_vars == b.comm_vars


def test_bincode_layout_of_the_computation_commitment():
    """ComputationCommitment = num_cons, num_vars, num_inputs, batch_size, num_ops, num_mem_cells (u64 LE each), then two
    Vec<CompressedRistretto> (u64 length + 32-byte items)  """
    built = O.build_point_add(*W.synth_point_add(4, infinity_every=3))
    sq, sp = W.tape_seeds()
    f = O.Flow(built, sq, sp, verify=False)
    u = lambda i: int.from_bytes(f.comm[8 * i:8 * i + 8], "little")
    num_cons, num_vars, num_inputs, batch, num_ops, num_mem = (u(i) for i in range(6))
    assert (num_cons, num_vars, num_inputs, batch) == (64, 64, 0, 3)
    assert num_ops == 64 and num_mem == 128  # next_pow2(nnz param 64); max(64, 2 * 64)
    n1 = u(6)
    off = 56 + 32 * n1
    n2 = int.from_bytes(f.comm[off:off + 8], "little")
    assert off + 8 + 32 * n2 == len(f.comm)
    assert n1 == 1 << ((num_ops * 16).bit_length() - 1) // 2 and n2 == 1 << ((num_mem * 2).bit_length() - 1) // 2
    # the proof starts with comm_vars: Vec of L compressed points
    L = int.from_bytes(f.proof[:8], "little")
    assert L == 8 and f.proof[8:8 + 32 * L] == f.comm_vars_para[:0] + O.Flow(built, sq, sp, verify=False).proof[8:8 + 32 * L]
    assert hashlib.sha256(f.proof).hexdigest() == GOLD["cases"][0]["proof_sha256"]
