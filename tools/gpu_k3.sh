#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
T=tools/tile_bench_dbg
$T 128 3 512 512 32 3 0 tile=2 > $O/k3_timeline_c3_fused.txt 2>&1
$T 128 3 512 512 32 3 1 tile=2 > $O/k3_timeline_c3_fwd.txt 2>&1
for f in $O/k3_timeline_c3_fused.txt $O/k3_timeline_c3_fwd.txt; do head -8 $f; done
