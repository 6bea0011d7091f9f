#!/bin/bash
# round-2 iteration 16: full suite, full bench, ncu for the cfg4 launch (now the tile kernel) and the S4 kernel, launch list of the "direct" step
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 > $O/k16_bench.json 2> $O/k16_bench.err; echo "bench rc=$?"; python tools/show_bench.py $O/k16_bench.json
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs none"
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 900 $NCU -k regex:warp_tile_kernel -s 3 -o $O/r2_cfg4 $B --workload cfg4 > $O/k16_ncu_cfg4.log 2>&1
timeout 600 $NCU -k regex:warp_persp -s 3 -o $O/r2_cfg3_persp python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs cfg3 > $O/k16_ncu_cfg3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_direct.csv $B --variant direct > $O/k16_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_dropin.csv $B --api dropin --variant direct > $O/k16_ncu_launch2.log 2>&1
