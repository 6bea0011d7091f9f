"""Per-CTA end times of the tile kernel inside the bench's own cfg2 step (development tool).

Needs the -DDMH_TILE_DEBUG build of the library (tools/build_tile_bench.sh -> tools/_dbg/libdmhomo.so):
    DMH_LIB=tools/_dbg/libdmhomo.so python tools/cta_spread.py [--workload cfg2] [--tuning k=v,...]
Runs the step eagerly over the rotating input sets, then reads the clocks the last tile launch left behind
(smid, start, end per CTA) and prints the spread: a balanced schedule has max / mean close to 1."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--variant", default="dlt")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--tuning", default="")
    ap.add_argument("--dump", default="")
    a = ap.parse_args()
    from dmhomo_b200 import _lib as L

    if a.tuning:
        L.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in a.tuning.split(","))})
    argv, sys.argv = sys.argv, sys.argv[:1]
    args = bench.parse()
    sys.argv = argv
    ctx = bench.Ctx(args)
    B = bench.WORKLOADS[a.workload]["B"] if a.workload != "cfg4" else 128
    st = bench.PairStep(ctx, a.workload, B, variant=a.variant)
    lib = L.lib()
    lib.dmh_tile_debug_dump.argtypes = [C.c_void_p, C.c_int]
    lib.dmh_tile_debug_dump.restype = C.c_int
    buf = np.zeros(1024 * 10, dtype=np.uint64)
    spreads, dumps = [], []
    for i in range(a.steps):
        st.step(i % len(st.sets))
        torch.cuda.synchronize()
        n = lib.dmh_tile_debug_dump(buf.ctypes.data, buf.nbytes)
        d = buf[: n * 10].reshape(n, 10).astype(np.float64)
        dumps.append(d.copy())
        t0 = d[:, 1].min()
        end = (d[:, 2] - t0) * 1e-3
        spreads.append((end.min(), end.mean(), end.max()))
    for i, s in enumerate(spreads):
        print(f"step {i:2d} set {i % len(st.sets)}: end_us min {s[0]:7.1f} mean {s[1]:7.1f} max {s[2]:7.1f}  max/mean {s[2] / s[1]:.3f}")
    if a.dump:
        # per-step CTA clocks + the homographies of every input set (offline cost model of the tile classes)
        out = {"cta": np.stack(dumps)}
        if a.workload == "cfg2" and a.variant == "dlt":
            from dmhomo_b200 import ops
            for k, sset in enumerate(st.sets):
                Hf, Hb = ops.basis_homography(st.basis, st.h, st.w, sset[2].detach(), sset[3].detach())
                out[f"H{k}"] = torch.stack([Hf, Hb]).cpu().numpy()
        np.savez(a.dump, **out)


if __name__ == "__main__":
    main()
