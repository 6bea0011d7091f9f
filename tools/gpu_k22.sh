#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
export DMH_LIB=tools/_dbg/libdmhomo.so
python tools/cta_spread.py --dump $O/k22_cfg2_cta.npz 2>&1 | grep step
