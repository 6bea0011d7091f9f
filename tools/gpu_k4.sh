#!/bin/bash
# round-2 kernel iteration 4: batched producer, serpentine tile order + direct dL/dH sums at C = 3
mkdir -p gpurun_out; O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
T=tools/tile_bench
{
echo "# cfg2 fused / forward"; $T 64 1 320 576 32 20; $T 64 1 320 576 32 20 1
echo "# cfg4 shape fused: tile=2 (tile kernel) vs tile=1 (scalar)"; $T 128 3 512 512 32 10 0 tile=2; $T 128 3 512 512 32 10 0 tile=1
echo "# cfg4 shape forward"; $T 128 3 512 512 32 10 1 tile=2; $T 128 3 512 512 32 10 1 tile=1
echo "# cfg5 frames forward"; $T 16 3 1080 1920 64 10 1 tile=2; $T 16 3 1080 1920 64 10 1 tile=1
echo "# cfg1 forward"; $T 16 1 360 640 32 20 1
} > $O/k4_tile_bench.txt 2>&1
cat $O/k4_tile_bench.txt
D=tools/tile_bench_dbg
for cfg in "64 1 320 576 32 3 0" "64 1 320 576 32 3 1" "128 3 512 512 32 3 0 tile=2" "128 3 512 512 32 3 1 tile=2"; do $D $cfg 2>&1 | head -8; done > $O/k4_timeline.txt
cat $O/k4_timeline.txt
NCU="ncu --set full --clock-control none --import-source on -k regex:warp_tile_kernel -s 3 -c 1 -f"
timeout 600 $NCU -o $O/k4_c3_fused $T 128 3 512 512 32 3 0 tile=2 > $O/k4_ncu1.log 2>&1
timeout 600 $NCU -o $O/k4_c3_fwd $T 128 3 512 512 32 3 1 tile=2 > $O/k4_ncu2.log 2>&1
