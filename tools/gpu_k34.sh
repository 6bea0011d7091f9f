#!/bin/bash
# round-2 final validation: suite, smoke, default bench + reference arm, corrected FFMA2 micro-benchmark
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 1200 python bench.py > $O/k34_bench.json 2> $O/k34_bench.err ) 2>&1 | tail -3; python tools/show_bench.py $O/k34_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/k34_bench_ref.json 2> /dev/null; echo "ref rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/k34_bench_ref.json').read().strip().splitlines()[-1]); print(d['impl'], d['value'], d['cpu_baseline'])"
tools/ubench_ffma2 > $O/r2_ubench_ffma2.txt 2>&1; cat $O/r2_ubench_ffma2.txt
