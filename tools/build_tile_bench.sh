#!/bin/bash
# Builds tools/tile_bench (against dmhomo_b200/libdmhomo.so) and tools/tile_bench_dbg (against a
# -DDMH_TILE_DEBUG build of the library under tools/_dbg/, per-CTA timeline dump).
set -e
cd "$(dirname "$0")/.."
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -std=c++17 -o tools/tile_bench tools/tile_bench.cu -Ldmhomo_b200 -ldmhomo -Xlinker -rpath -Xlinker '$ORIGIN/../dmhomo_b200'
mkdir -p tools/_dbg
for f in dmh_api dmh_warp dmh_warp_fast dmh_warp_tile dmh_dlt dmh_flow dmh_persp dmh_next; do
  mode=-dc; [ $f = dmh_warp_tile ] && mode=-c   # setmaxnreg needs whole-program compilation
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static -DDMH_TILE_DEBUG=${DMH_DBG_LEVEL:-1} $mode -o tools/_dbg/$f.o dmhomo_b200/csrc/$f.cu &
done
wait
nvcc $ARCH -shared -cudart static -o tools/_dbg/libdmhomo.so tools/_dbg/*.o
nvcc $ARCH -O3 -std=c++17 -o tools/tile_bench_dbg tools/tile_bench.cu -Ltools/_dbg -ldmhomo -Xlinker -rpath -Xlinker '$ORIGIN/_dbg'
