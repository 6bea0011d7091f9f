#!/bin/bash
# round-2 iteration 25: 24-consumer-warp geometry (tile_wide=1) for the gradient-free C = 1 launches: parity, then A/B
mkdir -p gpurun_out; O=gpurun_out
( DMH_TUNING=tile_wide=1 timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q ) > $O/pytest_gpu_wide.log 2>&1; tail -4 $O/pytest_gpu_wide.log
T=tools/tile_bench
{
echo "## default (16 warps, 64x64)"; $T 64 1 320 576 32 30 1; $T 16 1 360 640 32 30 1
echo "## tile_wide=1 (24 warps, 64x48)"; $T 64 1 320 576 32 30 1 tile_wide=1; $T 16 1 360 640 32 30 1 tile_wide=1
} > $O/k25_tile_bench.txt 2>&1
cat $O/k25_tile_bench.txt
for t in tile_wide=0 tile_wide=1; do
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg1,cfg2_dropin --tuning $t > $O/k25_bench.json 2> $O/k25_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/k25_bench.json').read().strip().splitlines()[-1]); c=d['configs']['cfg1']; print('$t cfg1', c['value'], c['ms_per_step'], c['kernel_ms'], c['roofline']['frac']); c=d['configs']['cfg2_dropin']; print('dropin', c['ms_per_step'], c['ms_per_step_from_flows'], c['kernel_ms'])"
done
