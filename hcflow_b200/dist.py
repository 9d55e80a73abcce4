"""Data-parallel plumbing: one process per GPU, batch sharded, no data-path collective.

The reference's only parallelism is batch sharding (codes/data/__init__.py:10-14,
codes/data/data_sampler.py:12-62) and every image is independent in eval mode, so the
inverse pass needs no exchange at all.  The single collective of the path is the batch
NLL: the reference takes ``nll.mean()`` over the batch (HCFlowNet_SR_arch.py:65); across
ranks that is one all-reduce of (sum nll_i, count).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of a batch of n_items; the first n_items % world ranks get one extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def batch_mean_nll(per_image_nll):
    """Mean of the per-image NLL over the GLOBAL batch: one all-reduce of (sum, count)."""
    s = torch.stack([per_image_nll.double().sum(), torch.tensor(float(per_image_nll.numel()),
                                                                dtype=torch.float64, device=per_image_nll.device)])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return (s[0] / s[1]).to(torch.float32)


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
