#!/bin/bash
# multi-GPU check (N = 4): bench.py under torchrun, all configs
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > $O/n4_bench.json 2> $O/n4_bench.err; echo "n4 rc=$?"
python tools/show_bench.py $O/n4_bench.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
