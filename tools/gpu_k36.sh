#!/bin/bash
# round-2 iteration 36: render_conditions (forked branches) - parity, cfg3 batch time
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "render or cfg3 or post_process" ) 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg3 > $O/k36_bench.json 2> $O/k36_bench.err; echo rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/k36_bench.json').read().strip().splitlines()[-1]); c=d['configs']['cfg3']; print('cfg3', c.get('us_per_batch'), c['ms_per_step'], c['kernel_ms'], c['roofline']['frac'], c['gpu_launches_per_step']); print([ (k['op'], round(k['us'],1)) for k in c['kernels']])"; tail -3 $O/k36_bench.err
