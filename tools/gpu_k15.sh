#!/bin/bash
# round-2 iteration 15: per-tile partial sums reduced by the drainer warp (C = 1): schedule A/B inside bench.py and in the harness
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
T="timeout 120 tools/tile_bench"
for knobs in "tile_dyn=100" "tile_dyn=100 tile_chunk=2" "tile_dyn=50" "tile_dyn=0"; do
echo "## $knobs"; $T 64 1 320 576 32 20 0 $knobs; $T 16 1 360 640 32 20 0 $knobs
done > $O/k15_tile_bench.txt 2>&1
cat $O/k15_tile_bench.txt
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/k15_$name.json 2> $O/k15_$name.err; echo "== $name rc=$?"; python tools/show_bench.py $O/k15_$name.json; }
run dyn100 --steps 30 --configs cfg1 --no-e2e --no-cpu-baseline
run dyn100c2 --steps 30 --configs none --no-e2e --no-cpu-baseline --tuning tile_chunk=2
run dyn100c4 --steps 30 --configs none --no-e2e --no-cpu-baseline --tuning tile_chunk=4
run dyn50 --steps 30 --configs none --no-e2e --no-cpu-baseline --tuning tile_dyn=50
run dyn0 --steps 30 --configs cfg1 --no-e2e --no-cpu-baseline --tuning tile_dyn=0
run pm1 --steps 30 --configs none --no-e2e --no-cpu-baseline --tuning tile_pair_major=1
tools/tile_bench_dbg 64 1 320 576 32 3 0 2>&1 | head -8
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:warp_persp -s 3 -o $O/k15_persp python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs cfg3 > $O/k15_ncu.log 2>&1
