#!/bin/bash
# multi-GPU check (N = 8): concurrent H2D probe, then bench.py under torchrun, all configs
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m > $O/n8_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe.py 2>/dev/null | grep H2D | tee $O/n8_h2d.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8_bench.json 2> $O/n8_bench.err; echo "n8 rc=$?"
python tools/show_bench.py $O/n8_bench.json; grep -v "torch.qr\|Q, R\|should be\|boolean parameter\|q, _ =\|W1017\|\*\*\*\*\|OMP_NUM" $O/n8_bench.err | tail -15
