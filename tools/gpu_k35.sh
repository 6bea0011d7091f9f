#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python bench.py --configs none --no-cpu-baseline > $O/k35_bench.json 2> $O/k35_bench.err; echo rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/k35_bench.json').read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['e2e']['variants'], indent=0)[:1200])"
python tools/h2d_probe.py 2>/dev/null | grep H2D
