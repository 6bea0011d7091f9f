#!/bin/bash
# round-2 iteration 11: explicit flow on the tile kernel; bench with e2e fixed
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
run() { name=$1; shift; timeout 900 python -X faulthandler bench.py "$@" > $O/k11_$name.json 2> $O/k11_$name.err; echo "== $name rc=$?"; python tools/show_bench.py $O/k11_$name.json; grep -v "torch.qr\|Q, R\|should be\|boolean parameter\|q, _ =" $O/k11_$name.err | tail -12; }
run full --steps 20
run static --steps 20 --configs none --no-e2e --no-cpu-baseline --tuning tile_dyn=0
run dyn25 --steps 20 --configs none --no-e2e --no-cpu-baseline --tuning tile_dyn=25,tile_chunk=1
run noflow --steps 20 --configs cfg2_direct,cfg2_dropin --no-e2e --no-cpu-baseline --tuning tile_flow=0
