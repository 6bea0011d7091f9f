#!/bin/bash
# round-2 iteration 42: dedicated many-channel forward warp kernel (pyramid levels)
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log | head -2; grep -n "^E " $O/pytest_gpu.log | head -5
timeout 900 python tools/bench_kernels.py 2>/dev/null | grep -E "pyramid|get_warp_flow" | cut -c1-150
