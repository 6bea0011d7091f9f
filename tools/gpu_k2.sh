#!/bin/bash
# round-2 kernel iteration 2: C = 3 tile kernel with the 96-word window pitch; ncu captures; light per-CTA timeline
mkdir -p gpurun_out; O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
T=tools/tile_bench
{
echo "# cfg2 fused / forward"; $T 64 1 320 576 32 20; $T 64 1 320 576 32 20 1
echo "# cfg4 shape fused: tile=2 (tile kernel) vs tile=1 (scalar)"; $T 128 3 512 512 32 10 0 tile=2; $T 128 3 512 512 32 10 0 tile=1
echo "# cfg4 shape forward"; $T 128 3 512 512 32 10 1 tile=2; $T 128 3 512 512 32 10 1 tile=1
echo "# cfg5 frames forward"; $T 16 3 1080 1920 64 10 1 tile=2; $T 16 3 1080 1920 64 10 1 tile=1
} > $O/k2_tile_bench.txt 2>&1
cat $O/k2_tile_bench.txt
tools/tile_bench_dbg 64 1 320 576 32 5 > $O/k2_timeline_c1.txt 2>&1; grep -A12 "per-CTA\|debug:" $O/k2_timeline_c1.txt | head -20
NCU="ncu --set full --clock-control none --import-source on -k regex:warp_tile_kernel -s 3 -c 1 -f"
timeout 600 $NCU -o $O/k2_c3_fused $T 128 3 512 512 32 3 0 tile=2 > $O/k2_ncu1.log 2>&1
timeout 600 $NCU -o $O/k2_c3_fwd $T 128 3 512 512 32 3 1 tile=2 > $O/k2_ncu2.log 2>&1
timeout 600 $NCU -o $O/k2_c1_fwd $T 64 1 320 576 32 3 1 > $O/k2_ncu3.log 2>&1
timeout 600 $NCU -o $O/k2_c1_fused $T 64 1 320 576 32 3 0 > $O/k2_ncu4.log 2>&1
ls -la $O/*.ncu-rep
