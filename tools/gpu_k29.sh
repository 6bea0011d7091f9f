#!/bin/bash
# round-2 iteration 29: channel groups on the lean kernel (C = 12 / 24 pyramid warps and any other C), bool masks as views
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 900 python tools/bench_kernels.py > $O/r2_kernels.txt 2> $O/k29_kernels.err; echo rc=$?; cat $O/r2_kernels.txt
