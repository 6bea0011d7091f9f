"""Concurrent host->device copy bandwidth of every rank of a torchrun job (development tool): what the box's PCIe /
host-memory fabric gives N GPUs at once, the ceiling of bench.py's end-to-end number.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mb = 256
    host = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    dev = torch.empty_like(host, device="cuda")
    stream = torch.cuda.Stream()
    res = []
    for size_mb in (24, 256):
        n = size_mb << 20
        with torch.cuda.stream(stream):
            for _ in range(3):
                dev[:n].copy_(host[:n], non_blocking=True)
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 40
            e0.record(stream)
            for _ in range(iters):
                dev[:n].copy_(host[:n], non_blocking=True)
            e1.record(stream)
            stream.synchronize()
        gbs = n * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
        res.append(gbs)
    t = torch.tensor(res, device="cuda", dtype=torch.float64)
    if world > 1:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
    else:
        allr = [t]
    if rank == 0:
        for i, size_mb in enumerate((24, 256)):
            per = [float(a[i]) for a in allr]
            print(f"H2D {size_mb} MB pinned copies, {world} rank(s) at once: per rank " + " ".join(f"{p:.1f}" for p in per) +
                  f" GB/s, aggregate {sum(per):.1f} GB/s", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
