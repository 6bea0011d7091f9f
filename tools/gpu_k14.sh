#!/bin/bash
# round-2 iteration 14: pair-major tile order (the two directions of a pair back to back: second use of every image / gradient line in the L2)
mkdir -p gpurun_out; O=gpurun_out
T="timeout 120 tools/tile_bench"
for knobs in "tile_pair_major=1" "tile_pair_major=0"; do
echo "## $knobs"
$T 64 1 320 576 32 20 0 $knobs; $T 64 1 320 576 32 20 1 $knobs
$T 128 3 512 512 32 10 0 tile=3 $knobs; $T 128 3 512 512 32 10 0 tile=3 tile_chunk=8 $knobs; $T 128 3 512 512 32 10 0 tile=3 tile_chunk=2 $knobs
$T 128 3 512 512 32 10 1 $knobs; $T 16 3 1080 1920 64 10 1 $knobs
$T 512 3 512 512 32 5 0 tile=3 $knobs
done > $O/k14_tile_bench.txt 2>&1
$T 128 3 512 512 32 10 0 tile=1 >> $O/k14_tile_bench.txt 2>&1
$T 512 3 512 512 32 5 0 tile=1 >> $O/k14_tile_bench.txt 2>&1
cat $O/k14_tile_bench.txt
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/k14_$name.json 2> $O/k14_$name.err; echo "== $name rc=$?"; python tools/show_bench.py $O/k14_$name.json; }
run pm1 --steps 30 --configs none --no-e2e --no-cpu-baseline
run pm0 --steps 30 --configs none --no-e2e --no-cpu-baseline --tuning tile_pair_major=0
run cfg4t3 --steps 6 --workload cfg4 --configs none --no-e2e --no-cpu-baseline --tuning tile=3
run cfg4t3pm0 --steps 6 --workload cfg4 --configs none --no-e2e --no-cpu-baseline --tuning tile=3,tile_pair_major=0
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:warp_tile_kernel -s 3 -o $O/k14_c3_fused_pm $T 128 3 512 512 32 3 0 tile=3 > $O/k14_ncu.log 2>&1
