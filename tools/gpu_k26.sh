#!/bin/bash
# round-2 iteration 26: 24-consumer-warp geometry with 64 x 96 tiles (RPT 8; 2 stages where a target tile is staged)
mkdir -p gpurun_out; O=gpurun_out
V=tools/_var/wide8
( DMH_LIB=$V/libdmhomo.so DMH_TUNING=tile_wide=1 timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py -x -q ) > $O/pytest_gpu_wide.log 2>&1; tail -3 $O/pytest_gpu_wide.log
T=tools/tile_bench
{
echo "## default (16 warps, 64x64)"; $T 64 1 320 576 32 30 1; $T 16 1 360 640 32 30 1
echo "## tile_wide=1 RPT 8 (24 warps, 64x96)"; LD_LIBRARY_PATH=$V $T 64 1 320 576 32 30 1 tile_wide=1; LD_LIBRARY_PATH=$V $T 16 1 360 640 32 30 1 tile_wide=1
} > $O/k26_tile_bench.txt 2>&1
cat $O/k26_tile_bench.txt
DMH_LIB=$V/libdmhomo.so timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg1 --tuning tile_wide=1 > $O/k26_bench.json 2> $O/k26_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/k26_bench.json').read().strip().splitlines()[-1]); c=d['configs']['cfg1']; print('wide8 cfg1', c['value'], c['ms_per_step'], c['kernel_ms'], c['roofline']['frac'])"
