"""One proof of a named vPIN shape on N GPUs with everything resident in HBM (api.prove_flow_resident on a distributed context),
once with the library's default partition (commitment rows + dealt product circuits) and once with only the rows sharded:
per-call times, prover phases of rank 0, digests compared across the ranks and with the golden digest when one is committed.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/prove_resident_dist.py L5 [reps]"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from vpin_b200 import api, workloads as W

tag = sys.argv[1] if len(sys.argv) > 1 else "E"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if (os.cpu_count() or 1) // world < 10:
    os.environ.setdefault("VPIN_HOST_HELPERS", "0")
ctx = api.Context(local)
ctx.init_distributed(rank, world, dist)
m, _ = W.SHAPES[tag]
weights, px, py = W.synth_point_mult(m)
sq, sp = W.tape_seeds()
gold = {}
try:
    gold = {(c["tag"], c["kind"]): c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_named.json")))["cases"]}
except OSError:
    pass
digests = []
for mode, name in ((-1, "default partition (rows + dealt product circuits)"), (0, "rows only")):
    ctx.set_shard_sumcheck(mode)
    for rep in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        tm = {}
        t = time.time(); got = api.prove_flow_resident(ctx, weights, px, py, sq, sp, timings=tm); dt = time.time() - t
        digests.append(hashlib.sha256(got["proof"]).hexdigest() + hashlib.sha256(got["comm"]).hexdigest())
        if rank == 0:
            print(f"{tag} on {world} GPUs, {name}, flow {rep}: {dt:.3f} s   " + "  ".join(f"{k} {v*1e3:.1f} ms" for k, v in tm.items()), flush=True)
            print("     prover phases (ms):", {k: round(x, 1) for k, x in ctx.phase_times().items() if not k.startswith("batched")}, flush=True)
        last = got
box = [None] * world
dist.all_gather_object(box, digests)
if rank == 0:
    same = all(b == box[0] for b in box) and len(set(digests)) == 1
    print(f"  every flow on every rank gave the same bytes: {same}")
    want = gold.get((tag, "point_mult"))
    if want:
        print("  golden digest:", "match" if digests[0][:64] == want["proof_sha256"] else "MISMATCH")
    else:
        import oracle_lib as O
        ok = O.verify(last["dims"], last["proof"], last["comm"], last["inputs"], last["comm_vars_para"], last["comm_vars_input"])
        print(f"  oracle my_lib_verify -> {ok}")
    assert same
del last, got
ctx.close()
dist.destroy_process_group()
