"""Per-kernel roofline table from an ncu launch list with a few metrics per launch:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,\
sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
      --clock-control none --csv --log-file gpurun_out/kernels.csv python scripts/prove_shape.py E 1
  python scripts/ncu_kernel_table.py gpurun_out/kernels.csv profiles/rN_kernels_E.csv

For every kernel: launches, total time, and — for its LONGEST launch — duration, DRAM GB/s against MEASURED_PEAKS.json, the
integer-multiply (fmaheavy) pipe utilisation, SM throughput, achieved occupancy, registers. ncu serialises launches and
the caches are cold, so totals are not step times; the per-launch figures of the long launches are what matters."""
import csv, json, os, re, sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    hbm = 6650.0
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]
ik, im, iu, iv, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
launches = OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(r[ii], {"kernel": re.sub(r"\(.*", "", r[ik]).replace("vpin::", "").replace("<unnamed>::", "").replace("void ", "")})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[im]] = v * scale
agg = OrderedDict()
for d in launches.values():
    a = agg.setdefault(d["kernel"], {"n": 0, "us": 0.0, "best": None})
    a["n"] += 1
    a["us"] += d["gpu__time_duration.sum"]
    if a["best"] is None or d["gpu__time_duration.sum"] > a["best"]["gpu__time_duration.sum"]:
        a["best"] = d
with open(dst, "w") as f:
    f.write("kernel,launches,total_ms,longest_us,longest_dram_GBps,longest_dram_frac_of_%.0f,longest_fmaheavy_pct,longest_sm_throughput_pct,"
            "longest_warps_active_pct,registers\n" % hbm)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        b = a["best"]
        us = b["gpu__time_duration.sum"]
        gbps = (b.get("dram__bytes_read.sum", 0) + b.get("dram__bytes_write.sum", 0)) / (us * 1e-6) / 1e9
        f.write(f"{k},{a['n']},{a['us'] / 1e3:.3f},{us:.1f},{gbps:.0f},{gbps / hbm:.3f},"
                f"{b.get('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):.1f},"
                f"{b.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f},"
                f"{b.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f},{b.get('launch__registers_per_thread', 0):.0f}\n")
print(open(dst).read())
