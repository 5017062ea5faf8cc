"""Data-parallel plumbing for the hot path: clips are independent units, so
ranks own disjoint clips and the only exchange is ONE all-reduce of the trainable gradients per
optimizer step (what Lightning's DDP strategy would do for the reference: README "Multi-GPU
Training", trainer.devices/strategy=ddp).  All 106 MemoryAttention gradients (5.9 M fp32 = 23.7 MB)
live in one flat buffer that autograd accumulates into in place, so the exchange is a single
NCCL all-reduce over NVLink with no bucket copies."""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


def shard_clips(num_clips: int, rank: int, world: int):
    """Clip indices owned by `rank` (rank r takes indices = r mod world, like DistributedSampler)."""
    return list(range(rank, num_clips, world))


class GradBucket:
    """One contiguous fp32 buffer holding every parameter gradient as a view."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def allreduce(self, world: Optional[int] = None, async_op: bool = False):
        world = world or (dist.get_world_size() if dist.is_initialized() else 1)
        if world <= 1:
            return None
        self.flat.div_(world)  # pre-divide: sum of means == mean, keeps magnitudes small
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)


def allreduce_gradients(model: torch.nn.Module, world: int):
    """Average gradients over ranks.  Uses the model's GradBucket if one is attached (no copies),
    else flattens on the fly."""
    bucket = getattr(model, "_sam2b200_grad_bucket", None)
    if bucket is not None:
        bucket.allreduce(world)
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads or world <= 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    flat.div_(world)
    dist.all_reduce(flat)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def attach_grad_bucket(model: torch.nn.Module) -> GradBucket:
    b = GradBucket(model.parameters())
    model._sam2b200_grad_bucket = b
    return b
