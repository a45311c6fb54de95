"""One named vPIN shape end to end with everything resident in HBM (api.prove_flow_resident): per-call times of the cold first
flow (generator tables built) and of warm flows, proof checked by the oracle's my_lib_verify and against the golden digest
when one is committed.   python scripts/prove_shape_resident.py L5 [reps]"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle_lib as O
from vpin_b200 import api, workloads as W

tag = sys.argv[1] if len(sys.argv) > 1 else "E"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m, _ = W.SHAPES[tag]
t = time.time(); weights, px, py = W.synth_point_mult(m); print(f"{tag}: m={m}, synthetic witness in {time.time()-t:.2f} s", flush=True)
ctx = api.Context(0)
sq, sp = W.tape_seeds()
first = None
for rep in range(reps):
    tm = {}
    t = time.time()
    got = api.prove_flow_resident(ctx, weights, px, py, sq, sp, timings=tm)
    dt = time.time() - t
    print(f"  flow {rep} ({'cold: generator tables built' if rep == 0 else 'warm'}): {dt:.3f} s   " + "  ".join(f"{k} {v*1e3:.1f} ms" for k, v in tm.items()), flush=True)
    print("     prover phases (ms):", {k: round(x, 1) for k, x in ctx.phase_times().items() if not k.startswith("batched")}, flush=True)
    print("     HBM in use: %.1f GB" % ((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9), flush=True)
    if first is None:
        first = got
        t = time.time()
        ok = O.verify(got["dims"], got["proof"], got["comm"], got["inputs"], got["comm_vars_para"], got["comm_vars_input"])
        print(f"  oracle my_lib_verify -> {ok} in {time.time()-t:.2f} s", flush=True)
        assert ok == 1
        try:
            gold = {(c["tag"], c["kind"]): c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_named.json")))["cases"]}
            want = gold.get((tag, "point_mult"))
            if want:
                print("  golden digest:", "match" if hashlib.sha256(got["proof"]).hexdigest() == want["proof_sha256"] else "MISMATCH", flush=True)
        except OSError:
            pass
    else:
        assert got["proof"] == first["proof"] and got["comm"] == first["comm"], "not deterministic"
