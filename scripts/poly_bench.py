"""HBM-roofline microbench of the F_l table kernels at large sizes (device resident): python scripts/poly_bench.py [log2_len]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vpin_b200 import api

ell = int(sys.argv[1]) if len(sys.argv) > 1 else 24
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ctx = api.Context(0)
dev = torch.device("cuda", 0)
lib = api.lib()
n = 1 << ell
g = torch.Generator(device="cpu").manual_seed(2)
def table():
    z = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
    z[:, 31] &= 0x0F
    return z.to(dev)
A, B, Cc, D = table(), table(), table(), table()
out = torch.zeros(32 * 8, dtype=torch.uint8, device=dev)
r = table()[:1].clone()
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
def timeit(name, fn, bytes_, reps=10):
    fn(); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); ctx.sync()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} n=2^{ell}: {ms:8.3f} ms  {bytes_/ms/1e6:8.1f} GB/s  {bytes_/ms/1e6/peak:6.3f} of measured HBM peak")
timeit("cubic round eval (4 tables)", lambda: ctx.check(lib.vpin_dev_cubic_round(ctx._h, p(A), p(B), p(Cc), p(D), C.c_uint64(n), p(out))), 4 * n * 32)
timeit("quad round eval (2 tables)", lambda: ctx.check(lib.vpin_dev_quad_round(ctx._h, p(A), p(B), C.c_uint64(n), p(out))), 2 * n * 32)
timeit("bind_top (1 table, r/w 1.5n)", lambda: ctx.check(lib.vpin_dev_bind_top(ctx._h, p(A), C.c_uint64(n), p(r))), 1.5 * n * 32)
rr = table()[:ell].clone()
timeit("eq_evals (write n)", lambda: ctx.check(lib.vpin_dev_eq_evals(ctx._h, p(rr), C.c_uint32(ell), p(B))), n * 32)
