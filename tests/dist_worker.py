"""Worker of tests/test_gpu_dist.py (one process per GPU, launched by torch.distributed.run): ONE proof of a named shape on
WORLD_SIZE GPUs - Hyrax rows sharded, then also the batched sumcheck rounds - must be byte-identical to the committed
single-prover golden digests on every rank."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from vpin_b200 import api, workloads as W


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_named.json")))
    cases = {(c["tag"], c["kind"]): c for c in gold["cases"]}
    sq, sp = (bytes.fromhex(s) for s in gold["tape_seeds"])
    ctx = api.Context(local)
    ctx.init_distributed(rank, world, dist)
    results = []
    for tag in sys.argv[1:] or ["conv3"]:
        m, n_add = W.SHAPES[tag]
        for kind in ("point_mult", "point_add"):
            if kind == "point_mult":
                dims, inst, vp, vi, v, inputs = api.point_mult(ctx, *W.synth_point_mult(m))
            else:
                dims, inst, vp, vi, v, inputs = api.point_addition(ctx, *W.synth_point_add(n_add))
            want = cases[(tag, kind)]
            for shard_rounds in (0, 1):
                ctx.set_shard_sumcheck(shard_rounds)
                got = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp)
                ok = (hashlib.sha256(got["proof"]).hexdigest() == want["proof_sha256"] and hashlib.sha256(got["comm"]).hexdigest() == want["comm_sha256"]
                      and hashlib.sha256(got["comm_vars"]).hexdigest() == want["comm_vars_sha256"])
                results.append((tag, kind, shard_rounds, ok))
                del got
            del inst
    box = [None] * world
    dist.all_gather_object(box, results)
    if rank == 0:
        bad = [(r, x) for r, res in enumerate(box) for x in res if not x[3]]
        print("DIST_RESULT", json.dumps({"world": world, "cases": len(results), "bad": bad}))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
