#!/bin/bash
# round-2 kernel iteration 5: L2 prefetch distance A/B (default 3 vs 0 vs 5), forward ILP 2
mkdir -p gpurun_out; O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
T=tools/tile_bench
run() { echo "## $1"; shift; "$@" 64 1 320 576 32 20; "$@" 64 1 320 576 32 20 1; "$@" 128 3 512 512 32 10 0 tile=2; "$@" 128 3 512 512 32 10 1 tile=2; "$@" 16 3 1080 1920 64 10 1 tile=2; "$@" 16 1 360 640 32 20 1; }
{
run "default (prefetch 3)" $T
run "prefetch 0" env LD_LIBRARY_PATH=tools/_var/pf0 $T
run "prefetch 5" env LD_LIBRARY_PATH=tools/_var/pf5 $T
run "fwd ilp 2" env LD_LIBRARY_PATH=tools/_var/ilp2 $T
} > $O/k5_tile_bench.txt 2>&1
cat $O/k5_tile_bench.txt
D=tools/tile_bench_dbg
for cfg in "64 1 320 576 32 3 0" "128 3 512 512 32 3 0 tile=2" "128 3 512 512 32 3 1 tile=2"; do $D $cfg 2>&1 | head -8; done > $O/k5_timeline.txt
cat $O/k5_timeline.txt
