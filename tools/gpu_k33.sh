#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log; grep -n "^E " $O/pytest_gpu.log | head -5
