#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 > $O/k18_bench.json 2> $O/k18_bench.err; echo "bench rc=$?"; python tools/show_bench.py $O/k18_bench.json
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_direct.csv $B --variant direct > $O/k18_ncu_launch.log 2>&1
