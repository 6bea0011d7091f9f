#!/bin/bash
# multi-GPU check (N = 2): bench.py under torchrun, all configs, both arms
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m > $O/n2_topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/n2_bench.json 2> $O/n2_bench.err; echo "n2 rc=$?"
python tools/show_bench.py $O/n2_bench.json; grep -v "torch.qr\|Q, R\|should be\|boolean parameter\|q, _ =\|W1017\|\*\*\*\*\|OMP_NUM" $O/n2_bench.err | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $O/n2_ref.json 2> $O/n2_ref.err; echo "ref rc=$?"; tail -c 600 $O/n2_ref.json
