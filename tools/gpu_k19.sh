#!/bin/bash
# round-2 iteration 19: mbarrier try_wait with / without a suspend-time hint (A/B), vectorised basis combine kernels
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
T=tools/tile_bench
run() { echo "## $1"; shift; "$@" 64 1 320 576 32 30; "$@" 64 1 320 576 32 30 1; "$@" 128 3 512 512 32 10 0; "$@" 128 3 512 512 32 10 1; "$@" 16 3 1080 1920 64 10 1; "$@" 16 1 360 640 32 30 1; }
{
run "hint 10 ms (default build)" $T
run "no hint" env LD_LIBRARY_PATH=tools/_var/nohint $T
run "hint 1 us" env LD_LIBRARY_PATH=tools/_var/hint1us $T
run "hint 10 ms again" $T
} > $O/k19_tile_bench.txt 2>&1
cat $O/k19_tile_bench.txt
timeout 900 python bench.py --steps 20 --no-cpu-baseline > $O/k19_bench.json 2> $O/k19_bench.err; echo "bench rc=$?"; python tools/show_bench.py $O/k19_bench.json
DMH_LIB=tools/_var/nohint/libdmhomo.so timeout 900 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg1,cfg5 > $O/k19_bench_nohint.json 2> $O/k19_bench_nohint.err; python tools/show_bench.py $O/k19_bench_nohint.json
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_direct.csv $B --variant direct > $O/k19_ncu_launch.log 2>&1
