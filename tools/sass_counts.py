#!/usr/bin/env python
"""profiles/r2_sass_tile.txt: per instantiation of dmh::warp_tile_kernel, the counts of the TMA / setmaxnreg / packed-fp32 /
mbarrier opcodes in the SASS of dmhomo_b200/libdmhomo.so (cuobjdump -sass).  Run after `make -C dmhomo_b200/csrc`."""
import collections, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "dmhomo_b200", "libdmhomo.so")], capture_output=True, text=True).stdout
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    m = re.search(r"warp_tile_kernelILi(\d+)ELi(\d+)ELb(\d)ELi(\d)E", name)
    if not m:
        continue
    mode, ct, s0, pk = [int(x) for x in m.groups()]
    ops = collections.Counter(re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f))
    cnt = lambda p: sum(v for k, v in ops.items() if k.startswith(p))
    rows.append((pk, ct, mode, s0, cnt("UTMALDG"), cnt("UTMASTG"), cnt("UTMAREDG"), cnt("UTMAPF"), cnt("USETMAXREG"), cnt("FFMA2"),
                 cnt("SYNCS"), cnt("REDG"), cnt("LDS"), sum(ops.values())))
rows.sort()
modes = {1: "OUT", 2: "LOSS", 3: "OUT+LOSS", 6: "LOSS+GRAD", 12: "GRAD+GOUT"}
out = ["# cuobjdump -sass dmhomo_b200/libdmhomo.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a), round 2: the instantiations of",
       "# dmh::warp_tile_kernel<MODE, C, START0, PK> (dmhomo_b200/csrc/dmh_warp_tile.cu) and the Blackwell / Hopper-class opcodes in them.",
       "# UTMALDG = TMA tensor load (cp.async.bulk.tensor), UTMASTG = TMA tensor store, UTMAREDG = TMA reduce-add (cp.reduce.async.bulk.tensor",
       "# .add), UTMAPF = TMA L2 prefetch (cp.async.bulk.prefetch.tensor), USETMAXREG = setmaxnreg (DEALLOC in the producer warpgroup, TRY_ALLOC",
       "# in the consumers), FFMA2 = packed fp32 FMA (fma.rn.f32x2), SYNCS = mbarrier operations, REDG = red.global.add.f32 (dL/dsrc scatter).",
       "# Regenerate: python tools/sass_counts.py", "",
       f"{'param':6s} {'C':>2s} {'mode':10s} {'start0':>6s} {'UTMALDG':>8s} {'UTMASTG':>8s} {'UTMAREDG':>9s} {'UTMAPF':>7s} {'USETMAXREG':>11s} {'FFMA2':>6s} {'SYNCS':>6s} {'REDG':>5s} {'LDS':>5s} {'total':>7s}"]
for r in rows:
    out.append(f"{('flow' if r[0] else 'H'):6s} {r[1]:>2d} {modes.get(r[2], str(r[2])):10s} {r[3]:>6d} {r[4]:>8d} {r[5]:>8d} {r[6]:>9d} {r[7]:>7d} {r[8]:>11d} {r[9]:>6d} {r[10]:>6d} {r[11]:>5d} {r[12]:>5d} {r[13]:>7d}")
tot = collections.Counter(re.findall(r"(UTMA[A-Z]+(?:\.[A-Z0-9]+)*|USETMAXREG\.[A-Z_]+\.[A-Z]+)", txt))
out += ["", "# whole library:"] + [f"#   {v:5d} {k}" for k, v in sorted(tot.items())]
open(os.path.join(ROOT, "profiles", "r2_sass_tile.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
