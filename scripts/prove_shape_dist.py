"""One proof of a named vPIN shape on N GPUs (one process per GPU, torchrun): every rank runs the same calls on the same
inputs, the rows of every Hyrax commitment are sharded across the ranks (NCCL all-gather of 32 B per row), rank 0 checks the
proof with the oracle's my_lib_verify:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/prove_shape_dist.py L5 [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from vpin_b200 import api, workloads as W

tag = sys.argv[1] if len(sys.argv) > 1 else "E"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if (os.cpu_count() or 1) // world < 10:
    os.environ.setdefault("VPIN_HOST_HELPERS", "0")
ctx = api.Context(local)
ctx.init_distributed(rank, world, dist)
m, _ = W.SHAPES[tag]
t = time.time(); dims, inst, vp, vi, v, inputs = api.point_mult(ctx, *W.synth_point_mult(m))
if rank == 0:
    print(f"{tag} on {world} GPUs: m={m} dims={dims} build {time.time()-t:.2f}s", flush=True)
sq, sp = W.tape_seeds()
first = None
for rep in range(reps):
    dist.barrier(); torch.cuda.synchronize()
    t = time.time(); got = api.prove_flow(ctx, dims, inst, vp, vi, v, inputs, sq, sp); dt = time.time() - t
    for k in ("decomm", "gens"):
        got.pop(k)
    if rank == 0:
        ph = ctx.phase_times()
        print(f"  prove_flow (gens+encode+commits+prove, host buffers) {dt:.3f}s  proof {len(got['proof'])} B", flush=True)
        print("   ", {k: round(x, 1) for k, x in ph.items() if not k.startswith('batched')}, flush=True)
    if first is None:
        first = got
        if rank == 0:
            import oracle_lib as O
            t = time.time()
            ok = O.verify(dims, got["proof"], got["comm"], inputs, got["comm_vars_para"], got["comm_vars_input"])
            print(f"  oracle my_lib_verify -> {ok} in {time.time()-t:.2f}s", flush=True)
            assert ok == 1
    else:
        assert got["proof"] == first["proof"], "not deterministic"
# every rank must hold the same proof bytes
h = torch.tensor(list(__import__("hashlib").sha256(first["proof"]).digest()), dtype=torch.uint8, device="cuda")
hs = [torch.empty_like(h) for _ in range(world)]
dist.all_gather(hs, h)
assert all(bool((x == hs[0]).all()) for x in hs), "ranks disagree on the proof"
if rank == 0:
    print(f"  all {world} ranks hold the same proof bytes")
del inst, first, got
ctx.close()
dist.destroy_process_group()
