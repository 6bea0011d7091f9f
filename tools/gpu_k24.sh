#!/bin/bash
# round-2 iteration 24: two row pairs in flight in the C = 1 training fast body (A/B)
mkdir -p gpurun_out; O=gpurun_out
T=tools/tile_bench
{
echo "## default"; $T 64 1 320 576 32 30; $T 128 3 512 512 32 10
echo "## grad ILP 2"; LD_LIBRARY_PATH=tools/_var/gilp2 $T 64 1 320 576 32 30; LD_LIBRARY_PATH=tools/_var/gilp2 $T 128 3 512 512 32 10
} > $O/k24_tile_bench.txt 2>&1
cat $O/k24_tile_bench.txt
for v in "" tools/_var/gilp2/libdmhomo.so; do
DMH_LIB=$v timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e --configs cfg4 > $O/k24_bench.json 2> $O/k24_bench.err; python tools/show_bench.py $O/k24_bench.json
python -c "
import json; d=json.loads(open('gpurun_out/k24_bench.json').read().strip().splitlines()[-1]); c=d['configs']['cfg4']; print('cfg4', c['value'], c['kernel_ms'], c['roofline']['frac'])"
done
