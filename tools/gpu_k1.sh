#!/bin/bash
# round-2 kernel iteration 1: parity of the restructured tile kernel (C = 1 and C = 3), then A/B timings through the C-ABI harness
mkdir -p gpurun_out; O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
T=tools/tile_bench
{
echo "# cfg2 fused / forward"; $T 64 1 320 576 32 20; $T 64 1 320 576 32 20 1
echo "# cfg4 shape fused: tile=2 (tile kernel) vs tile=1 (scalar)"; $T 128 3 512 512 32 10 0 tile=2; $T 128 3 512 512 32 10 0 tile=1
echo "# cfg4 shape forward"; $T 128 3 512 512 32 10 1 tile=2; $T 128 3 512 512 32 10 1 tile=1
echo "# cfg5 frames forward"; $T 16 3 1080 1920 64 10 1 tile=2; $T 16 3 1080 1920 64 10 1 tile=1
echo "# cfg1 forward"; $T 16 1 360 640 32 20 1
} > $O/k1_tile_bench.txt 2>&1
cat $O/k1_tile_bench.txt
timeout 300 python bench.py --no-cpu-baseline --steps 30 > $O/k1_bench_cfg2.json 2>$O/k1_bench_cfg2.err; python tools/show_bench.py $O/k1_bench_cfg2.json
timeout 300 python bench.py --no-cpu-baseline --workload cfg4 --steps 10 > $O/k1_bench_cfg4.json 2>$O/k1_bench_cfg4.err; python tools/show_bench.py $O/k1_bench_cfg4.json
DMH_TUNING=tile=1 timeout 300 python bench.py --no-cpu-baseline --workload cfg4 --steps 10 > $O/k1_bench_cfg4_scalar.json 2>/dev/null; python tools/show_bench.py $O/k1_bench_cfg4_scalar.json
