#!/bin/bash
# round-2 closing run: suite, smoke, default bench, launch list of the "direct" step, per-kernel table
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log | head -2
timeout 900 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 1200 python bench.py > $O/k40_bench.json 2> $O/k40_bench.err ) 2>&1 | grep real; python tools/show_bench.py $O/k40_bench.json
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_direct.csv $B --variant direct > $O/k40_ncu_launch2.log 2>&1
timeout 900 python tools/bench_kernels.py > $O/r2_kernels.txt 2> /dev/null; cat $O/r2_kernels.txt | cut -c1-150
