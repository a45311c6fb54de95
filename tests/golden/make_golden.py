"""Generates tests/golden/golden_flows.json: digests of everything the driver flow emits (computation commitment, the
three witness commitments, bincode proof) for small seeded instances, produced by the CPU oracle (oracle/liboracle.so).

The reference itself (Rust) cannot run in this image, so these vectors pin the ORACLE against regressions and let the GPU
box check the CUDA path against committed bytes without re-deriving them; what pins the oracle to the reference is
tests/test_oracle_primitives.py (reference KATs, RFC 9496, merlin vector, libsodium).

    python tests/golden/make_golden.py          # rewrites golden_flows.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers as H  # noqa: E402
import oracle_lib as O  # noqa: E402
from vpin_b200 import workloads as W  # noqa: E402

CASES = [
    ("point_add", dict(n=4, infinity_every=3)),
    ("point_add", dict(n=16, infinity_every=0)),
    ("point_mult", dict(m=7)),
    ("synthetic", dict(num_cons=16, num_vars=16, num_inputs=3)),
    ("synthetic", dict(num_cons=37, num_vars=50, num_inputs=2)),
    ("synthetic", dict(num_cons=1, num_vars=2, num_inputs=1)),
]


def build(kind, kw):
    if kind == "point_add":
        return O.build_point_add(*W.synth_point_add(kw["n"], infinity_every=kw["infinity_every"]))
    if kind == "point_mult":
        return O.build_point_mult(*W.synth_point_mult(kw["m"]))
    A, B, Cm, vp, vi, v, inputs = H.synthetic_r1cs(kw["num_cons"], kw["num_vars"], kw["num_inputs"], seed=kw["num_cons"])
    nnz = max(len(A), 2) if kw["num_cons"] > 1 else 2
    return O.build_custom(kw["num_cons"], kw["num_vars"], kw["num_inputs"], nnz, A, B, Cm, vp, vi, v, inputs)


def digest(b):
    return hashlib.sha256(b).hexdigest()


def run(kind, kw):
    built = build(kind, kw)
    sq, sp = W.tape_seeds()
    f = O.Flow(built, sq, sp, verify=True)
    assert f.verified
    return dict(kind=kind, args=kw, dims=list(built.dims), proof_len=len(f.proof), proof_sha256=digest(f.proof), comm_sha256=digest(f.comm),
                comm_vars_para_sha256=digest(f.comm_vars_para), comm_vars_input_sha256=digest(f.comm_vars_input),
                comm_vars_sha256=digest(f.comm_vars), proof_head=f.proof[:64].hex())


if __name__ == "__main__":
    out = dict(tape_seeds=[s.hex() for s in W.tape_seeds()], transcript_label="snark_example", cases=[run(k, kw) for k, kw in CASES])
    json.dump(out, open(os.path.join(HERE, "golden_flows.json"), "w"), indent=1)
    print("wrote", len(out["cases"]), "cases")
