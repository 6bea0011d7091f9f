#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python tools/bench_kernels.py > $O/r2_kernels.txt 2> $O/k45.err; echo rc=$?; cut -c1-150 $O/r2_kernels.txt; tail -2 $O/k45.err
