"""Multi-GPU plumbing: the path shards by batch (independent pairs / frames); the only
exchange is one tiny all-reduce(sum) of loss / count scalars per step,
mirroring the reference's scalar `accelerator.gather` (hem_evaluate.py:132-151)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of a batch of n for `rank` (accelerate split_batches=True,
    HEM/train.py:179): the first n % world ranks take one extra element."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank, world, dim=0):
    lo, hi = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def all_reduce_sums(vec):
    """In-place sum of a small vector of partial sums (loss numerators, counts, error sums) on
    the current stream; the caller divides afterwards so the result equals the single-GPU mean
    over the global batch."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def global_mean_loss(local_loss_mean, local_count, out=None):
    """Combine per-rank means over unequal shards: sum(mean_r * n_r) / sum(n_r)."""
    v = torch.stack([local_loss_mean.double() * float(local_count),
                     torch.tensor(float(local_count), dtype=torch.float64, device=local_loss_mean.device)])
    all_reduce_sums(v)
    return (v[0] / v[1]).float()
