#!/bin/bash
# Quick iteration: tile-kernel parity tests, cfg2 bench (with and without the interior body), harness numbers, one ncu capture.
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_quick.log
timeout 300 python bench.py --no-cpu-baseline --steps 30 > $O/q_bench_cfg2.json 2>$O/q_bench_cfg2.err; python tools/show_bench.py $O/q_bench_cfg2.json
DMH_TUNING=tile_interior=0 timeout 300 python bench.py --no-cpu-baseline --steps 30 > $O/q_bench_cfg2_noint.json 2>/dev/null; python tools/show_bench.py $O/q_bench_cfg2_noint.json
for extra in "$@"; do
  env $extra timeout 300 python bench.py --no-cpu-baseline --steps 30 > $O/q_tmp.json 2>/dev/null; echo "$extra"; python tools/show_bench.py $O/q_tmp.json
done
tools/tile_bench 64 1 320 576 32 20 > $O/q_tile_bench.txt 2>&1
tools/tile_bench 64 1 320 576 32 20 1 >> $O/q_tile_bench.txt 2>&1
tools/tile_bench 128 3 512 512 32 10 tile=2 >> $O/q_tile_bench.txt 2>&1
tools/tile_bench 16 3 1080 1920 64 10 1 >> $O/q_tile_bench.txt 2>&1
tools/tile_bench 16 3 1080 1920 64 10 1 tile=2 >> $O/q_tile_bench.txt 2>&1
cat $O/q_tile_bench.txt
[ -x tools/tile_bench_dbg ] && tools/tile_bench_dbg 64 1 320 576 32 5 > $O/q_tile_timeline.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp_tile_kernel -s 6 -c 1 -f -o $O/q_tile_full \
  python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > $O/q_ncu_full.log 2>&1
