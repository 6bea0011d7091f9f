#!/bin/bash
# round-2 iteration 27 (evidence run of the final build): full suite, default bench, ncu --set full per workload, launch lists
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > $O/k27_bench.json 2> $O/k27_bench.err; echo "bench rc=$?"; python tools/show_bench.py $O/k27_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/k27_bench_ref.json 2> $O/k27_bench_ref.err; echo "ref rc=$?"; tail -c 600 $O/k27_bench_ref.json
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs none"
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:warp_tile_kernel -s 6 -o $O/r2_cfg2 $B > $O/k27_ncu_cfg2.log 2>&1
timeout 600 $NCU -k regex:warp_tile_kernel -s 6 -o $O/r2_cfg1 $B --workload cfg1 > $O/k27_ncu_cfg1.log 2>&1
timeout 600 $NCU -k regex:warp_tile_kernel -s 6 -o $O/r2_cfg2_direct $B --variant direct > $O/k27_ncu_cfg2d.log 2>&1
timeout 900 $NCU -k regex:warp_tile_kernel -s 3 -o $O/r2_cfg4 $B --workload cfg4 > $O/k27_ncu_cfg4.log 2>&1
timeout 900 $NCU -k regex:warp_tile_kernel -s 8 -o $O/r2_cfg5 $B --workload cfg5 > $O/k27_ncu_cfg5.log 2>&1
timeout 600 $NCU -k regex:warp_persp -s 3 -o $O/r2_cfg3_persp python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --configs cfg3 > $O/k27_ncu_cfg3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_step.csv $B > $O/k27_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_direct.csv $B --variant direct > $O/k27_ncu_launch2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_dropin.csv $B --api dropin --variant direct > $O/k27_ncu_launch3.log 2>&1
ls -la $O/r2_*.ncu-rep
