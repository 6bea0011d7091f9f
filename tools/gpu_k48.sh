#!/bin/bash
# round-2 iteration 48: compute-sanitizer memcheck over the kernels added late in the round
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py -x -q -k "remap or multichannel or pyramid or basis_combine or render_conditions or warp_perspective or soft_mask_and_flow" > $O/r2_sanitizer.txt 2>&1; echo "memcheck rc=$?"; tail -6 $O/r2_sanitizer.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tile.py -x -q -k "flow_forward_backward and smooth" > $O/r2_sanitizer_race.txt 2>&1; echo "racecheck rc=$?"; tail -4 $O/r2_sanitizer_race.txt
