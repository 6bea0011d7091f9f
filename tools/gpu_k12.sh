#!/bin/bash
# round-2 iteration 12: tests, bench A/B, ncu captures per workload (DRAM traffic of the dominant kernel), launch list
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/k12_$name.json 2> $O/k12_$name.err; echo "== $name rc=$?"; python tools/show_bench.py $O/k12_$name.json; }
run base --steps 30 --configs cfg1 --no-e2e --no-cpu-baseline
run cfg1static --steps 30 --configs cfg1 --no-e2e --no-cpu-baseline --tuning tile_dyn=0
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs none"
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:warp_tile_kernel -s 6 -o $O/r2_cfg2 $B > $O/k12_ncu_cfg2.log 2>&1
timeout 600 $NCU -k regex:warp_tile_kernel -s 6 -o $O/r2_cfg1 $B --workload cfg1 > $O/k12_ncu_cfg1.log 2>&1
timeout 600 $NCU -k regex:warp_tile_kernel -s 6 -o $O/r2_cfg2_direct $B --variant direct > $O/k12_ncu_cfg2d.log 2>&1
timeout 900 $NCU -k regex:warp_fast_kernel -s 3 -o $O/r2_cfg4 $B --workload cfg4 > $O/k12_ncu_cfg4.log 2>&1
timeout 900 $NCU -k regex:warp_tile_kernel -s 8 -o $O/r2_cfg5 $B --workload cfg5 > $O/k12_ncu_cfg5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_step.csv $B > $O/k12_ncu_launch.log 2>&1
ls -la $O/r2_*.ncu-rep
