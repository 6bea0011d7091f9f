#!/bin/bash
# round-2 iteration 41: the headline step through ops.basis_warp_loss (forked branches) vs the two-call composition
mkdir -p gpurun_out; O=gpurun_out
for v in 1 0 1 0; do DMH_BENCH_ONEOP=$v timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs none > $O/k41_b.json 2>$O/k41_b.err; python -c "
import json; d=json.loads(open('gpurun_out/k41_b.json').read().strip().splitlines()[-1]); print('oneop=$v', round(d['value'],1), 'step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches'], 'loss', d['run']['loss'])" || tail -5 $O/k41_b.err; done
