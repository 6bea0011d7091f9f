#!/bin/bash
# round-2 iteration 37: transposing warp reduction in the per-sample flush of the tile kernel
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log | head -2
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg4,cfg2_direct > $O/k37_bench.json 2>/dev/null; python tools/show_bench.py $O/k37_bench.json; python -c "
import json; d=json.loads(open('gpurun_out/k37_bench.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],4), round(v['kernel_ms'],4), round(v['roofline']['frac'],3))"
