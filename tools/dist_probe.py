"""Probe: latency of the path's only collective (a 16-byte all-reduce) eagerly and inside a CUDA graph.
torchrun --nproc-per-node N tools/dist_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
v = torch.ones(2, device=dev, dtype=torch.float64)
s = torch.cuda.Stream(dev)
def log(*a):
    if rank == 0: print(*a, flush=True)
with torch.cuda.stream(s):
    for _ in range(5): dist.all_reduce(v)
    s.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): dist.all_reduce(v)
    s.synchronize()
    log(f"eager all_reduce(16 B) x{world}: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us each")
    if "--graph" in sys.argv:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            v.mul_(0.5); dist.all_reduce(v)
        log("captured")
        for _ in range(5): g.replay()
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(200): g.replay()
        s.synchronize()
        log(f"graph replay (mul + all_reduce): {(time.perf_counter() - t0) / 200 * 1e6:.1f} us each")
dist.barrier(); dist.destroy_process_group(); log("done")
