#!/bin/bash
# round-2 kernel iteration 8: guided dynamic schedule, share / run-length sweep; then the whole GPU suite and bench.py
mkdir -p gpurun_out; O=gpurun_out
T="timeout 120 tools/tile_bench"
for knobs in "tile_dyn=100" "tile_dyn=100 tile_chunk=5" "tile_dyn=100 tile_chunk=3" "tile_dyn=50" "tile_dyn=25" "tile_dyn=15 tile_chunk=1"; do
echo "## $knobs"
$T 64 1 320 576 32 20 0 $knobs; $T 64 1 320 576 32 20 1 $knobs
$T 128 3 512 512 32 10 0 tile=3 $knobs; $T 128 3 512 512 32 10 1 $knobs
$T 16 3 1080 1920 64 10 1 $knobs; $T 16 1 360 640 32 20 1 $knobs
done > $O/k8_tile_bench.txt 2>&1
cat $O/k8_tile_bench.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 > $O/k8_bench.json 2> $O/k8_bench.err; echo "bench rc=$?"; tail -c 3000 $O/k8_bench.json; tail -5 $O/k8_bench.err
