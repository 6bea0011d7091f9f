#!/bin/bash
# usage: tools/tune.sh "768 800 816 ..." [workload] -> kernel_ms / ms_per_step per DMH_TUNE value (needs a DMH_TUNE_BUILD library)
WL=${2:-cfg2}
for t in $1; do
  DMH_TUNE=$t python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$WL TUNE=$t', 'kernel_ms=%.4f'%d['roofline']['kernel_ms'], 'ms_per_step=%.4f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'])"
done
