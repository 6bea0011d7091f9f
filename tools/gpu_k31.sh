#!/bin/bash
# round-2 iteration 31: golden test through the reference's names, H2D probe on one GPU
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_golden.py -x -q ) 2>&1 | tail -3
python tools/h2d_probe.py 2>&1 | tail -3 | tee $O/k31_h2d_n1.txt
