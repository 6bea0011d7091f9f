#!/usr/bin/env python
"""profiles/traffic.json from the per-workload `ncu --set full` captures of tools/gpu_k12.sh:
    python tools/ncu_traffic.py cfg2=gpurun_out/r2_cfg2.ncu-rep cfg1=... [--commit <hash>]
dram__bytes_read.sum + dram__bytes_write.sum of the one captured launch of the workload's dominant kernel (bytes),
read by bench.py into roofline.traffic / roofline.dram_gbs."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
meta = {"unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, one launch, ncu --set full --clock-control none)"}
UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
args = sys.argv[1:]
if "--commit" in args:
    del args[args.index("--commit"):args.index("--commit") + 2]
for arg in args:
    key, _, rep = arg.partition("=")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print(key, "unreadable", rep)
        continue
    h, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(name)
        tot += float(vals[i].replace(",", "")) * UNITS[units[i]]
    dur_i = h.index("gpu__time_duration.sum")
    dur = float(vals[dur_i].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[dur_i].replace("second", "s").replace("nsecond", "ns"), 1.0)
    out[key] = int(tot)
    meta[key + "_kernel"] = vals[h.index("Kernel Name")] if "Kernel Name" in h else ""
    meta[key + "_duration_under_ncu"] = f"{vals[dur_i]} {units[dur_i]}"
    print(f"{key}: {tot / 1e6:.1f} MB per launch, {vals[dur_i]} {units[dur_i]} under ncu")
if "--commit" in sys.argv:
    meta["commit"] = sys.argv[sys.argv.index("--commit") + 1]
# output pixels of the captured launch (one GPU): bench.py scales the bytes when a rank's launch covers fewer (N > 1, strong scaling)
PIXELS = {"cfg1": 2 * 16 * 360 * 640, "cfg2": 2 * 64 * 320 * 576, "cfg2_direct": 2 * 64 * 320 * 576, "cfg3": 25 * 256 * 256,
          "cfg4": 2 * 4096 * 512 * 512, "cfg5": 512 * 1080 * 1920}
out["_pixels"] = {k: PIXELS[k] for k in out if k in PIXELS}
out["_meta"] = meta
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
