#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for args in "3 smooth" "0 smooth" "3 noise" "0 noise"; do
echo "## channels / flow: $args"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/k47_pyr.csv python tools/time_pyramid_warp.py $args > /dev/null 2>&1
python - <<'P'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/k47_pyr.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
d=collections.OrderedDict()
for r in rows[1:]:
    if 'warp' in r[ki]: d.setdefault((r[ki][20:75], r[gi]), []).append(float(r[vi])/1000)
for k,v in d.items(): print(k[0], k[1], ' '.join(f"{x:.1f}" for x in v), 'us')
P
done
