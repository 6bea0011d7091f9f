#!/bin/bash
# round-2 iteration 32: per-sample sums flushed on sample change (direct, transposing reduction): parity, then the dynamic share again
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
sweep() { for t in "$@"; do timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs none --tuning $t > $O/k32_b.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/k32_b.json').read().strip().splitlines()[-1]); print('$t', 'step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3))"; done; }
sweep tile_dyn=0 tile_dyn=10,tile_chunk=1 tile_dyn=20,tile_chunk=1 tile_dyn=30,tile_chunk=1 tile_dyn=20,tile_chunk=2 tile_dyn=30,tile_chunk=2 tile_dyn=50,tile_chunk=2 tile_dyn=100,tile_chunk=4 tile_dyn=100,tile_chunk=8 tile_dyn=0 2>&1 | tee $O/k32_sweep.txt
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg1,cfg4,cfg5,cfg2_direct > $O/k32_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/k32_bench.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],4), round(v['kernel_ms'],4), round(v['roofline']['frac'],3))"
