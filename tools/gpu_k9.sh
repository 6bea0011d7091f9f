#!/bin/bash
# round-2 iteration 9: whole GPU suite + bench.py (all configs)
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 > $O/k9_bench.json 2> $O/k9_bench.err; echo "bench rc=$?"; tail -c 6000 $O/k9_bench.json; tail -5 $O/k9_bench.err
