"""Uniform full-width Hyrax commitment microbench (device resident): python scripts/msm_bench.py [ell] [reps] [label]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vpin_b200 import api

ell = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 252
ctx = api.Context(0)
dev = torch.device("cuda", 0)
n = 1 << ell
g = torch.Generator(device="cpu").manual_seed(1)
z = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
z[:, 31] &= 0x0F
if bits < 252:
    nb = bits // 8
    z[:, nb:] = 0
d = z.to(dev)
torch.cuda.synchronize()
api.dev_to_mont(ctx, d, n, d)
out = torch.empty(32 * (1 << (ell // 2)), dtype=torch.uint8, device=dev)
lib = api.lib()
label = b"gens_r1cs_eval"
def run():
    ctx.check(lib.vpin_dev_hyrax_commit(ctx._h, label, C.c_void_p(d.data_ptr()), C.c_uint64(n), None, C.c_void_p(out.data_ptr())))
run(); ctx.sync()
peak = ctx.imad_peak()
ctx.profile_enable(True, 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps): run()
e1.record(stream); ctx.sync()
sec = e0.elapsed_time(e1) / 1e3 / reps
prof, madds = ctx.profile_read()
print(f"ell={ell} bits={bits}: {sec*1e3:.3f} ms/commit, {n/sec/1e6:.1f} Mpoints/s, algorithmic frac {n*8064/sec/peak:.4f}, peak {peak/1e12:.2f} TMAC/s")
for k, v in prof.items():
    extra = f" executed frac {madds*504/(v['ms']*1e-3)/peak:.4f}" if k == "msm_accumulate" else ""
    print(f"  {k:18s} {v['ms']/reps:9.3f} ms{extra}")
print("checksum", bytes(out[:16].cpu().numpy()).hex())
