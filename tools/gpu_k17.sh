#!/bin/bash
# round-2 iteration 17: explicit-flow fast body (per-pair vote), basis_combine backward in sample groups
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/k17_$name.json 2> $O/k17_$name.err; echo "== $name rc=$?"; python tools/show_bench.py $O/k17_$name.json; }
run direct --steps 20 --configs cfg2_direct,cfg2_dropin --no-e2e --no-cpu-baseline
run direct_noint --steps 20 --configs cfg2_direct,cfg2_dropin --no-e2e --no-cpu-baseline --tuning tile_interior=1
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_direct.csv $B --variant direct > $O/k17_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_dropin.csv $B --api dropin --variant direct > $O/k17_ncu_launch2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:basis_combine_bwd -s 2 -o $O/k17_bcb $B --variant direct > $O/k17_ncu_bcb.log 2>&1
timeout 600 $NCU -k regex:warp_tile_kernel -s 6 -o $O/r2_cfg2_direct $B --variant direct > $O/k17_ncu_cfg2d.log 2>&1
