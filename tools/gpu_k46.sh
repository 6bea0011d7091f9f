#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; grep -n "passed\|failed" $O/pytest_gpu.log | tail -1; grep -n "^E " $O/pytest_gpu.log | head -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/k46_pyr.csv python tools/time_pyramid_warp.py 2>/dev/null | sort | uniq -c
python - <<'P'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/k46_pyr.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
for r in rows[1:]:
    if 'dmh' in r[ki]: print(r[ki][:70], r[gi], r[vi])
P
