#!/bin/bash
# round-2 iteration 28: per-kernel roofline table of the secondary launches with the final build (profiles/r2_kernels.txt)
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python tools/bench_kernels.py > $O/r2_kernels.txt 2> $O/k28_kernels.err; echo rc=$?; cat $O/r2_kernels.txt; tail -3 $O/k28_kernels.err
