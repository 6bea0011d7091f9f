#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + per-opcode executed instructions per warp-pixel and stall mix.
usage: ncu_summary.py report.ncu-rep <pixels per launch> [--list MIN]"""
import collections, csv, io, subprocess, sys
rep, npx = sys.argv[1], float(sys.argv[2]) / 32
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Executed Ipc Active", "DRAM Throughput",
        "L1/TEX Hit Rate", "L2 Hit Rate", "Executed Instructions", "Memory Throughput", "Avg. Active Threads", "Not Predicated",
        "Issue Slots Busy", "Mem Pipes Busy")
for line in det.splitlines():
    if any(k in line for k in keys) or line.strip().startswith("void ") or "warp_" in line[:60]:
        print(line.rstrip()[:150])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
if len(rows) > 2:
    h = rows[0]
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum"):
        if name in h:
            print(name, rows[2][h.index(name)], rows[1][h.index(name)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ci, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ci]) for r in data)
print(f"executed warp-instr {tot}  = {tot / npx:.1f} per warp-pixel")
by, bys = collections.Counter(), collections.Counter()
for r in data:
    toks = r[si].split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    by[op] += int(r[ci]); bys[op] += int(r[ss])
tots = sum(bys.values())
print("  ".join(f"{k}:{v / npx:.1f}({100 * bys[k] / tots:.0f}%)" for k, v in by.most_common(26)))
st = []
for name in hdr:
    if name.startswith("stall_") and "Not Issued" not in name:
        s = sum(int(r[hdr.index(name)]) for r in data)
        if s > 0.02 * tots: st.append(f"{name[6:]}:{100 * s / tots:.0f}%")
print("stalls:", " ".join(st))
if "--list" in sys.argv:
    mn = float(sys.argv[sys.argv.index("--list") + 1])
    for i, r in enumerate(data):
        n = int(r[ci]) / npx
        if n >= mn: print(f"{i:5d} {n:5.2f} {int(r[ss]):6d}  {r[si].strip()}")
