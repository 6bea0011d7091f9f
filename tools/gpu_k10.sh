#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
run() { name=$1; shift; timeout 600 python -X faulthandler bench.py "$@" > $O/k10_$name.json 2> $O/k10_$name.err; echo "== $name rc=$?"; tail -c 1500 $O/k10_$name.json; grep -v "torch.qr\|Q, R\|should be\|boolean parameter\|q, _ =" $O/k10_$name.err | tail -25; }
run base --steps 5 --configs none --no-e2e --no-cpu-baseline
run e2e --steps 5 --configs none --no-cpu-baseline
run cfg1 --steps 5 --configs cfg1 --no-e2e --no-cpu-baseline
run cfg2v --steps 5 --configs cfg2_direct,cfg2_dropin --no-e2e --no-cpu-baseline
run cfg3 --steps 5 --configs cfg3 --no-e2e --no-cpu-baseline
run cfg4 --steps 5 --configs cfg4 --no-e2e --no-cpu-baseline
run cfg5 --steps 5 --configs cfg5 --no-e2e --no-cpu-baseline
