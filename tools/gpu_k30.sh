#!/bin/bash
# round-2 iteration 30: suite again (mass-based bound of the flow test), margins of the fp64-yardstick checks, three repeats of the scatter-heavy tests
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
DMH_TEST_REPORT=1 timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -s 2>&1 | grep "\[margin\]" > $O/k30_margins.txt; cat $O/k30_margins.txt
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py tests/test_gpu_next.py -x -q 2>&1 | tail -1; done
