#!/bin/bash
# round-2 iteration 21: per-CTA end-time spread of the tile kernel inside the bench's own cfg2 / cfg1 steps
mkdir -p gpurun_out; O=gpurun_out
export DMH_LIB=tools/_dbg/libdmhomo.so
python tools/cta_spread.py --dump $O/k21_cfg2_cta.npy > $O/k21_spread.txt 2>&1
python tools/cta_spread.py --variant direct >> $O/k21_spread.txt 2>&1
python tools/cta_spread.py --workload cfg4 --steps 4 >> $O/k21_spread.txt 2>&1
cat $O/k21_spread.txt
