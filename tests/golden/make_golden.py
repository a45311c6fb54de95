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




# ---- BASELINE.json's named shapes at full size (golden_named.json) ----------------------------------------------------------
# The CPU oracle needs seconds (conv3) to minutes (CNN E) per instance, far too long for a test, so the digests are generated
# once here and committed; tests/test_gpu_golden.py::test_named_shape_matches_golden compares the CUDA flow with them byte
# for byte (sha256 of the proof and of the four commitments, proof length, first 64 proof bytes).
#     python tests/golden/make_golden.py named [tags...]      # default tags: conv3 conv5 conv7 A E
NAMED_DEFAULT = ["conv3", "conv5", "conv7", "A", "E"]


def run_named(tag, which):
    import time
    m, n_add = W.SHAPES[tag]
    if which == "mult":
        built, args = O.build_point_mult(*W.synth_point_mult(m)), dict(m=m)
    else:
        built, args = O.build_point_add(*W.synth_point_add(n_add)), dict(n=n_add, infinity_every=0)
    sq, sp = W.tape_seeds()
    t0 = time.time()
    f = O.Flow(built, sq, sp, verify=True)
    assert f.verified
    print(f"  {tag}/{which}: dims {built.dims}, proof {len(f.proof)} B, oracle flow {time.time() - t0:.1f} s", flush=True)
    return dict(tag=tag, kind="point_" + which, args=args, dims=list(built.dims), proof_len=len(f.proof), proof_sha256=digest(f.proof),
                comm_sha256=digest(f.comm), comm_vars_para_sha256=digest(f.comm_vars_para),
                comm_vars_input_sha256=digest(f.comm_vars_input), comm_vars_sha256=digest(f.comm_vars), proof_head=f.proof[:64].hex())


def main_named(tags):
    path = os.path.join(HERE, "golden_named.json")
    out = json.load(open(path)) if os.path.exists(path) else dict(tape_seeds=[s.hex() for s in W.tape_seeds()],
                                                                   transcript_label="snark_example", cases=[])
    for tag in tags:
        for which in ("mult", "add"):
            out["cases"] = [c for c in out["cases"] if not (c["tag"] == tag and c["kind"] == "point_" + which)]
            out["cases"].append(run_named(tag, which))
            json.dump(out, open(path, "w"), indent=1)
    print("wrote", len(out["cases"]), "named cases")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "named":
        main_named(sys.argv[2:] or NAMED_DEFAULT)
    else:
        out = dict(tape_seeds=[s.hex() for s in W.tape_seeds()], transcript_label="snark_example", cases=[run(k, kw) for k, kw in CASES])
        json.dump(out, open(os.path.join(HERE, "golden_flows.json"), "w"), indent=1)
        print("wrote", len(out["cases"]), "cases")
