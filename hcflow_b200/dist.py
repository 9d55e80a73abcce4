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


class GradientReducer:
    """Gradient all-reduce of a data-parallel training step, overlapped with backward (third bullet; the
    reference gets it from DistributedDataParallel, HCFlow_SR_model.py:33-36 / optimize_parameters :195-218).

    Parameters are packed, in REVERSE registration order (backward reaches the last layers first), into flat buckets of
    about ``bucket_bytes``.  A post-accumulate-grad hook copies each finished ``.grad`` into its bucket; the moment a
    bucket is complete its all-reduce is issued asynchronously (NCCL runs it on its own stream while the backward
    kernels of the earlier layers keep going on the compute stream).  ``finish()`` -- call it after ``backward()``,
    before ``optimizer.step()`` -- flushes buckets left incomplete (parameters that got no gradient count as zero),
    waits, divides by the world size (DDP's mean) and writes the result back into every ``.grad``.

    The gradients of the whole SR x4 net are 23 M floats: two or three buckets in flight at the default size.
    """

    def __init__(self, params, bucket_bytes=32 << 20, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.average = average
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets = []          # [flat, [(param, offset, numel)], n_ready, work]
        self._slot = {}
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or cur[0].dtype != p.dtype or cur[0].device != p.device):
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.launched_early = 0    # buckets whose all-reduce was issued from inside backward (statistics for tests)

    def _close(self, plist):
        total = sum(p.numel() for p in plist)
        flat = torch.zeros(total, dtype=plist[0].dtype, device=plist[0].device)
        entries, off = [], 0
        for p in plist:
            entries.append((p, off, p.numel()))
            self._slot[p] = (len(self.buckets), len(entries) - 1)
            off += p.numel()
        self.buckets.append({"flat": flat, "entries": entries, "ready": set(), "work": None})

    def _on_grad(self, p):
        bi, ei = self._slot[p]
        b = self.buckets[bi]
        _, off, n = b["entries"][ei]
        b["flat"][off:off + n].copy_(p.grad.reshape(-1))
        b["ready"].add(ei)
        if len(b["ready"]) == len(b["entries"]) and b["work"] is None:
            self._launch(b)
            self.launched_early += 1

    def _launch(self, b):
        if self.world > 1:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, async_op=True)
        else:
            b["work"] = True

    def finish(self):
        """Wait for every bucket, average, write back.  Returns the number of buckets reduced."""
        for b in self.buckets:
            if b["work"] is None:
                for ei, (p, off, n) in enumerate(b["entries"]):
                    if ei not in b["ready"]:
                        b["flat"][off:off + n].zero_()
                self._launch(b)
        for b in self.buckets:
            if b["work"] is not True:
                b["work"].wait()
            if self.average and self.world > 1:
                b["flat"].div_(self.world)
            for ei, (p, off, n) in enumerate(b["entries"]):
                if p.grad is not None:
                    p.grad.copy_(b["flat"][off:off + n].view_as(p.grad))
                elif self.world > 1:
                    p.grad = b["flat"][off:off + n].view_as(p).clone()    # another rank had a gradient for it
            b["ready"].clear()
            b["work"] = None
        return len(self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
