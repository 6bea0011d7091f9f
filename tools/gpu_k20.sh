#!/bin/bash
# round-2 iteration 20: per-CTA end times of the cfg2 training launch (static split), two runs: SM-bound or tile-mix-bound spread?
mkdir -p gpurun_out; O=gpurun_out
D=tools/tile_bench_dbg
$D 64 1 320 576 32 5 0 > $O/k20_cta_run1.txt 2>&1
$D 64 1 320 576 32 5 0 > $O/k20_cta_run2.txt 2>&1
$D 64 1 320 576 32 5 0 tile_dyn=100 tile_chunk=1 > $O/k20_cta_dyn.txt 2>&1
$D 64 1 320 576 32 5 1 > $O/k20_cta_fwd.txt 2>&1
head -3 $O/k20_cta_run1.txt $O/k20_cta_dyn.txt
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k basis ) 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg2_direct > $O/k20_bench.json 2> $O/k20_bench.err; python tools/show_bench.py $O/k20_bench.json
