"""End-to-end parity: the CUDA prover behind the C ABI must emit the same commitments and the same bincode proof bytes as
the CPU oracle under the same tape seeds, and the oracle's my_lib_verify must accept the CUDA proof."""
import pytest

import helpers as H
import oracle_lib as O
from vpin_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _check(ctx, built, dims, inst, vp, vi, v, inputs):
    from vpin_b200 import api

    sq, sp = W.tape_seeds()
    ref = O.Flow(built, sq, sp, verify=True)
    assert ref.verified
    got = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
    assert got["comm"] == ref.comm, "SNARK::encode commitment differs"
    assert got["comm_vars_para"] == ref.comm_vars_para
    assert got["comm_vars_input"] == ref.comm_vars_input
    assert got["comm_vars"] == ref.comm_vars
    if got["proof"] != ref.proof:
        n = min(len(got["proof"]), len(ref.proof))
        first = next((i for i in range(n) if got["proof"][i] != ref.proof[i]), n)
        raise AssertionError(f"proof bytes differ at offset {first} (lengths {len(got['proof'])} vs {len(ref.proof)})")
    assert O.verify(dims, got["proof"], got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"]) == 1
    # a corrupted proof must be rejected
    bad = bytearray(got["proof"])
    bad[len(bad) // 2] ^= 1
    assert O.verify(dims, bytes(bad), got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"]) != 1


@pytest.mark.parametrize("n,inf", [(4, 3), (16, 0)])
def test_point_add_flow(ctx, n, inf):
    from vpin_b200 import api

    px, py, rx, ry, rz = W.synth_point_add(n, infinity_every=inf)
    built = O.build_point_add(px, py, rx, ry, rz)
    dims, inst, vp, vi, v, inputs = api.point_addition(ctx, px, py, rx, ry, rz)
    assert dims == built.dims
    A, B, Cm, ovp, ovi, ov, oin = built.arrays()
    assert (vp, vi, v) == (ovp, ovi, ov)
    _check(ctx, built, dims, inst, vp, vi, v, inputs)


def test_point_mult_flow_m7(ctx):
    """smallest point-mult count whose hard-coded nnz table rounds to the true padded nnz"""
    from vpin_b200 import api

    weights, px, py = W.synth_point_mult(7)
    built = O.build_point_mult(weights, px, py)
    dims, inst, vp, vi, v, inputs = api.point_mult(ctx, weights, px, py)
    assert dims == built.dims
    A, B, Cm, ovp, ovi, ov, oin = built.arrays()
    assert (vp, vi, v, inputs) == (ovp, ovi, ov, oin)
    assert inst.is_sat(v, inputs)
    _check(ctx, built, dims, inst, vp, vi, v, inputs)


@pytest.mark.parametrize("m", [1, 7, 33])
def test_point_mult_device_builder(ctx, m):
    """vpin_build_point_mult_device (witness expansion and COO emission on the device, assignments left in HBM in Montgomery
    form) against the oracle's restatement of point_mult.rs: every assignment byte, every COO triple in the reference's
    order, zero padding, and satisfiability."""
    import numpy as np
    import torch
    from vpin_b200 import api

    weights, px, py = W.synth_point_mult(m, seed=W.SEED + 17 * m)
    if m > 1:  # edge weights: zero, one, all 128 bits set
        weights = [0, 1, (1 << 128) - 1] + list(weights[3:])
    built = O.build_point_mult(weights, px, py)
    dims, inst, d_para, d_input, d_vars, inputs, padded = api.point_mult_device(ctx, weights, px, py)
    assert dims == built.dims
    A, B, Cm, ovp, ovi, ov, oin = built.arrays()
    assert inputs == oin
    nv = dims[1]
    for d, want in ((d_para, ovp), (d_input, ovi), (d_vars, ov)):
        canon = torch.empty_like(d)
        api.dev_from_mont(ctx, d, padded, canon)
        ctx.sync()
        got = canon.cpu().numpy().tobytes()
        assert got[: 32 * nv] == want
        assert got[32 * nv:] == bytes(32 * (padded - nv))
    for got, want in zip(inst.export_coo(nv), (A, B, Cm)):
        assert np.array_equal(got, want)
    assert inst.is_sat(ov, oin)
    # the host-buffer builder is the same device code plus a download
    dims2, inst2, vp, vi, v, inputs2 = api.point_mult(ctx, weights, px, py)
    assert (dims2, vp, vi, v, inputs2) == (dims, ovp, ovi, ov, oin)


@pytest.mark.parametrize("num_cons,num_vars,num_inputs", [(1, 2, 1), (16, 16, 3), (64, 32, 5), (1024, 1024, 10), (37, 50, 2)])
def test_synthetic_r1cs_flow(ctx, num_cons, num_vars, num_inputs):
    """shapes of the reference's own round-trip tests (Spartan/src/lib.rs:615-774, r1csproof.rs:586-619) incl. padding"""
    from vpin_b200 import api

    A, B, Cm, vp, vi, v, inputs = H.synthetic_r1cs(num_cons, num_vars, num_inputs, seed=num_cons)
    nnz = max(len(A), 2) if num_cons > 1 else 2
    built = O.build_custom(num_cons, num_vars, num_inputs, nnz, A, B, Cm, vp, vi, v, inputs)
    inst = api.Instance(ctx, num_cons, num_vars, num_inputs, A, B, Cm)
    _check(ctx, built, (num_cons, num_vars, num_inputs, nnz), inst, vp, vi, v, inputs)


def test_padded_constraints_instance_of_the_reference(ctx):
    """The instance of Spartan's own `test_padded_constraints` (Spartan/src/lib.rs:694-774): ONE constraint a^2 = z - 13 - b over
    zero variables and three public inputs (z = 16, a = 1, b = 2) — the constraint count is padded to 2, the variable count to
    4 (max(num_vars, num_inputs + 1) rounded up), and the witness is empty. Through vPIN's flow, against the oracle."""
    import numpy as np
    from vpin_b200 import api

    L = O.L_ORDER
    num_cons, num_vars, num_inputs, nnz = 1, 0, 3, 3
    def ent(rows):
        a = np.zeros(len(rows), O.COO_DTYPE)
        for k, (r, c, val) in enumerate(rows):
            a[k] = (r, c, np.frombuffer(O.le32(val % L), dtype=np.uint8))
        return a
    A = ent([(0, num_vars + 2, 1)])
    B = ent([(0, num_vars + 2, 1)])
    Cm = ent([(0, num_vars + 1, 1), (0, num_vars, -13), (0, num_vars + 3, -1)])
    inputs = O.le32(16) + O.le32(1) + O.le32(2)
    built = O.build_custom(num_cons, num_vars, num_inputs, nnz, A, B, Cm, b"", b"", b"", inputs)
    inst = api.Instance(ctx, num_cons, num_vars, num_inputs, A, B, Cm)
    assert inst.is_sat(b"", inputs)
    assert not inst.is_sat(b"", O.le32(17) + O.le32(1) + O.le32(2))
    _check(ctx, built, (num_cons, num_vars, num_inputs, nnz), inst, b"", b"", b"", inputs)


def test_unsatisfied_witness_is_reported(ctx):
    from vpin_b200 import api

    A, B, Cm, vp, vi, v, inputs = H.synthetic_r1cs(16, 16, 2, seed=3)
    inst = api.Instance(ctx, 16, 16, 2, A, B, Cm)
    bad = bytearray(v)
    bad[0] ^= 1
    assert inst.is_sat(v, inputs) and not inst.is_sat(bytes(bad), inputs)


def test_derefs_row_half_on_the_side_stream(ctx, monkeypatch):
    """The row half of the derefs commitment is committed on a second, low-priority stream while the second sumcheck runs
    (prover.cu, SideScope; the default on one GPU, VPIN_DEREFS_EARLY=0 keeps everything on one stream); the proof bytes must
    not depend on it."""
    from vpin_b200 import api

    weights, px, py = W.synth_point_mult(7)
    dims, inst, vp, vi, v, inputs = api.point_mult(ctx, weights, px, py)
    sq, sp = W.tape_seeds()
    early = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
    monkeypatch.setenv("VPIN_DEREFS_EARLY", "0")
    base = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
    monkeypatch.setenv("VPIN_DEREFS_EARLY", "1")
    again = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
    assert early["proof"] == base["proof"] == again["proof"] and early["comm"] == base["comm"]


def test_null_seeds_draw_fresh_randomness(ctx):
    """NULL tape seeds (the production default): the library draws both init_randomness scalars from the OS like the reference's
    OsRng (SP/random.rs:16-18) - two proofs of one witness then differ in every blinded byte, and both verify."""
    from vpin_b200 import api

    px, py, rx, ry, rz = W.synth_point_add(8, infinity_every=3)
    dims, inst, vp, vi, v, inputs = api.point_addition(ctx, px, py, rx, ry, rz)
    a = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, None, None)
    b = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, None, None)
    assert a["comm"] == b["comm"]                      # SNARK::encode is deterministic (no blinds)
    assert a["comm_vars"] != b["comm_vars"] and a["proof"] != b["proof"] and len(a["proof"]) == len(b["proof"])
    for got in (a, b):
        assert O.verify(dims, got["proof"], got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"]) == 1


def test_two_contexts_prove_concurrently():
    """A network's independent instances are proved from two host threads with one context each (bench.py does this for the
    point-add / point-mult pair, the point-mult context on the urgent stream priority): both proofs must be the oracle's."""
    from concurrent.futures import ThreadPoolExecutor
    from vpin_b200 import api

    sq, sp = W.tape_seeds()
    jobs = []
    px, py, rx, ry, rz = W.synth_point_add(24, infinity_every=7)
    jobs.append(("add", False, O.build_point_add(px, py, rx, ry, rz), lambda c: api.point_addition(c, px, py, rx, ry, rz)))
    weights, mx, my = W.synth_point_mult(7)  # the smallest count whose hard-coded nnz parameter fits
    jobs.append(("mult", True, O.build_point_mult(weights, mx, my), lambda c: api.point_mult(c, weights, mx, my)))

    def run(job):
        _, urgent, _, build = job
        c = api.Context(0, high_priority=urgent)
        dims, inst, vp, vi, v, inputs = build(c)
        res = []
        for _ in range(2):
            o = api.prove_flow(c, dims, inst, vp, vi, v, inputs, sq, sp)
            res.append((o["comm"], o["comm_vars"], o["proof"]))
            del o  # the generator / decommitment handles go before the context that owns their stream
        del inst
        c.close()
        return res

    with ThreadPoolExecutor(max_workers=2) as pool:
        results = list(pool.map(run, jobs))
    for (name, _, built, _), outs in zip(jobs, results):
        ref = O.Flow(built, sq, sp, verify=True)
        assert ref.verified
        for comm, comm_vars, proof in outs:
            assert (comm, comm_vars) == (ref.comm, ref.comm_vars), name
            assert proof == ref.proof, name


def test_encode_on_a_background_context_under_the_commitments_and_the_proof(ctx):
    """What bench.py's step does on one GPU: a helper thread runs SNARK::encode on a background context - the dense tables while
    the first context commits to the assignments, the computation commitment's MSMs while it proves (my_lib_prove reads the
    tables only). The same computation commitment as the one-call encode, and the proof made with that decommitment is the
    oracle's."""
    from concurrent.futures import ThreadPoolExecutor
    from vpin_b200 import api

    sq, sp = W.tape_seeds()
    weights, mx, my = W.synth_point_mult(7)
    built = O.build_point_mult(weights, mx, my)
    ref = O.Flow(built, sq, sp, verify=True)
    dims, inst, vp, vi, v, inputs = api.point_mult(ctx, weights, mx, my)
    gens = api.SNARKGens(ctx, *dims)
    aux = api.Context(0, background=True)
    with ThreadPoolExecutor(max_workers=1) as pool:
        for _ in range(2):
            f_tab = pool.submit(api.encode_tables, inst, gens, aux)
            f_comm = pool.submit(lambda: api.encode_commit(f_tab.result(), gens, aux))
            tape = api.RandomTape(b"\x02", sq)
            p_para, p_input, p_vars = inst.pad(vp), inst.pad(vi), inst.pad(v)
            c_para, b_para = api.dense_mlpoly_commit(ctx, gens, p_para, tape)
            c_input, b_input = api.dense_mlpoly_commit(ctx, gens, p_input, tape)
            c_vars, b_vars = api.my_dense_mlpoly_commit(ctx, gens, p_vars, b_para, b_input)
            combined = ctx.commitments_add(c_para, c_input)
            decomm = f_tab.result()
            proof = api.my_lib_prove(inst, decomm, p_vars, inputs, gens, b"snark_example", combined, b_vars, sp)
            comm = f_comm.result()
            assert comm == ref.comm and c_vars == ref.comm_vars
            assert proof == ref.proof
            del decomm
        # the one-call form on a second, ordinary context
        comm2, decomm2 = api.SNARK.encode(inst, gens, aux)
        assert comm2 == ref.comm
        del decomm2
        # Instance::new and SNARK::encode both on the second context, the proof on the first (bench.py's e2e step)
        inst2 = api.Instance(aux, dims[0], dims[1], dims[2], *inst.export_coo(dims[1]))
        comm3, decomm3 = api.SNARK.encode(inst2, gens, aux)
        proof3 = api.my_lib_prove(inst2, decomm3, p_vars, inputs, gens, b"snark_example", combined, b_vars, sp, ctx=ctx)
        assert comm3 == ref.comm and proof3 == ref.proof
        del decomm3, inst2
    del gens, inst
    aux.close()


def _mailbox_worker(env):
    import json
    import os
    import subprocess
    import sys

    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "mailbox_worker.py")], env=e,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_prelaunch_on_and_off_give_the_same_bytes():
    """Rounds enqueued before their challenge exists (the kernel waits for the host's post in mapped memory) against challenges
    passed as kernel parameters: the same proof bytes, and both equal to the committed oracle digest of the m = 7 flow."""
    import json
    import os

    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_flows.json")))
    want = [c for c in gold["cases"] if c["kind"] == "point_mult" and c["args"] == {"m": 7}]
    on = _mailbox_worker({"VPIN_PRELAUNCH_Q": "1048576"})
    off = _mailbox_worker({"VPIN_PRELAUNCH_Q": "0"})
    assert [r["proof"] for r in on] == [r["proof"] for r in off] and on[0]["proof"] == on[1]["proof"]
    assert on[0]["comm"] == off[0]["comm"]
    if want:
        assert on[0]["proof"] == want[0]["proof_sha256"]


def test_prelaunch_lost_post_times_out_and_the_proof_is_redone():
    """A host that never posts a challenge (VPIN_TEST_DROP_POST loses the 40th post): the waiting kernel gives up after
    VPIN_MAILBOX_TIMEOUT_MS, the context stops pre-launching and redoes the proof - same bytes, and the context keeps working."""
    ok = _mailbox_worker({})
    lost = _mailbox_worker({"VPIN_TEST_DROP_POST": "40", "VPIN_MAILBOX_TIMEOUT_MS": "200"})
    assert [r["proof"] for r in lost] == [r["proof"] for r in ok]
    assert lost[0]["s"] < 30


@pytest.mark.parametrize("tag", ["conv3", "conv5", "A"])
def test_named_vpin_shapes_verify(ctx, tag):
    """BASELINE.json's named shapes at full size (point-mult instance): too large for the CPU prover inside a test, so the
    check is the size-independent one — the oracle's restatement of my_lib_verify (O(sqrt n)) must accept the CUDA proof,
    reject a corrupted one, and the combined commitment must be the row-wise sum (proof_point_add.rs:69-80)."""
    from vpin_b200 import api

    m, _ = W.SHAPES[tag]
    weights, px, py = W.synth_point_mult(m)
    dims, inst, vp, vi, v, inputs = api.point_mult(ctx, weights, px, py)
    sq, sp = W.tape_seeds()
    got = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
    assert ctx.commitments_add(got["comm_vars_para"], got["comm_vars_input"]) == got["comm_vars"]
    assert O.verify(dims, got["proof"], got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"]) == 1
    bad = bytearray(got["proof"])
    bad[len(bad) // 3] ^= 0x10
    assert O.verify(dims, bytes(bad), got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"]) != 1
    # determinism: the same seeds give the same bytes
    again = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
    assert again["proof"] == got["proof"] and again["comm"] == got["comm"]
