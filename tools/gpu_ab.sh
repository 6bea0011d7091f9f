#!/bin/bash
# A/B of library variants / env knobs through bench.py (kernel_ms of the dominant launch): tools/gpu_ab.sh "ENV=.. ENV=.." ...
mkdir -p gpurun_out; O=gpurun_out; : > $O/ab.txt
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --steps 40 > $O/ab_tmp.json 2>$O/ab_tmp.err || tail -3 $O/ab_tmp.err
  echo "== $cfg" | tee -a $O/ab.txt; python tools/show_bench.py $O/ab_tmp.json | tee -a $O/ab.txt
done
