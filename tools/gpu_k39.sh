#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py -x -q ) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log; grep -n "^E " $O/pytest_gpu.log | head -8
