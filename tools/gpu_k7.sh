#!/bin/bash
# round-2 kernel iteration 7: dynamic tail of the tile schedule (classifier claims), share sweep
mkdir -p gpurun_out; O=gpurun_out
T="timeout 120 tools/tile_bench"
for dyn in 15 0 8 25 40 100; do
echo "## tile_dyn=$dyn"
$T 64 1 320 576 32 20 0 tile_dyn=$dyn; $T 64 1 320 576 32 20 1 tile_dyn=$dyn
$T 128 3 512 512 32 10 0 tile=2 tile_dyn=$dyn; $T 128 3 512 512 32 10 1 tile=2 tile_dyn=$dyn
$T 16 3 1080 1920 64 10 1 tile=2 tile_dyn=$dyn; $T 16 1 360 640 32 20 1 tile_dyn=$dyn
done > $O/k7_tile_bench.txt 2>&1
cat $O/k7_tile_bench.txt
( time timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
D="timeout 120 tools/tile_bench_dbg"
for cfg in "64 1 320 576 32 3 0" "64 1 320 576 32 3 1" "128 3 512 512 32 3 1 tile=2"; do $D $cfg 2>&1 | head -8; done > $O/k7_timeline.txt
cat $O/k7_timeline.txt
