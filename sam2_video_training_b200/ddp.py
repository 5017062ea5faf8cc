"""Data-parallel plumbing for the hot path: clips are independent units, so
ranks own disjoint clips and the only exchange is ONE all-reduce of the trainable gradients per
optimizer step (what Lightning's DDP strategy would do for the reference: README "Multi-GPU
Training", trainer.devices/strategy=ddp).  All 106 MemoryAttention gradients (5.9 M fp32 = 23.7 MB)
live in one flat buffer that autograd accumulates into in place, so the exchange is a single
NCCL all-reduce over NVLink with no bucket copies."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_clips(num_clips: int, rank: int, world: int):
    """Clip indices owned by `rank` (rank r takes indices = r mod world, like DistributedSampler)."""
    return list(range(rank, num_clips, world))


class GradBucket:
    """One contiguous fp32 buffer holding every parameter gradient as a view.

    `order` (optional) lists the parameters in the order they should be laid out, so that gradients one kernel
    produces together are adjacent (MemoryAttention puts each layer's q/k/v self-attention weights -- and biases --
    next to each other: one [768, 256] weight-gradient GEMM writes all three).  Every offset is a multiple of 64
    floats (256 bytes), which keeps every view aligned for vectorised kernels and cuBLAS."""

    ALIGN = 64

    def __init__(self, params: Iterable[torch.nn.Parameter], order: Optional[List[torch.nn.Parameter]] = None):
        self.params = [p for p in (order if order is not None else params) if p.requires_grad]
        if order is not None:
            listed = {id(p) for p in self.params}
            self.params += [p for p in params if p.requires_grad and id(p) not in listed]
        dev = self.params[0].device
        self.offset = {}
        off = 0
        for p in self.params:
            self.offset[id(p)] = off
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.attach()

    def attach(self):
        """(Re-)install the views as the parameters' .grad.  A gradient that already exists elsewhere (accumulated before
        the bucket was attached, or after ``zero_grad(set_to_none=True)`` detached the views) is carried over."""
        with torch.no_grad():
            for p in self.params:
                o = self.offset[id(p)]
                view = self.flat[o:o + p.numel()].view_as(p)
                if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                p.grad = view

    def owns_all(self) -> bool:
        return all(self.owns(p) for p in self.params)

    def span(self, plist, shape):
        """One view over the gradients of `plist` if they are laid out back to back (no padding), else None."""
        o = self.offset.get(id(plist[0]))
        if o is None:
            return None
        exp = o
        for p in plist:
            if self.offset.get(id(p)) != exp or p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * exp:
                return None
            exp += p.numel()
        return self.flat[o:exp].view(shape)

    def owns(self, p) -> bool:
        o = self.offset.get(id(p))
        return (o is not None and p.grad is not None and p.grad.dtype == torch.float32
                and p.grad.data_ptr() == self.flat.data_ptr() + 4 * o)

    def zero(self):
        self.flat.zero_()

    def allreduce(self, world: Optional[int] = None, async_op: bool = False):
        world = world or (dist.get_world_size() if dist.is_initialized() else 1)
        if world <= 1:
            return None
        self.flat.div_(world)  # pre-divide: sum of means == mean, keeps magnitudes small
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)


def allreduce_gradients_async(model: torch.nn.Module, world: int):
    """Start the gradient all-reduce NOW (on NCCL's own stream, after everything enqueued so far on the current stream) and
    return a handle; ``handle.wait()`` makes the current stream wait for it.  Called right after the last backward that
    touches the parameters, so that the collective overlaps whatever follows on the compute stream before the optimizer
    (in the hot path: the mask-loss forward / backward, which has no parameters).  Returns None when there is nothing to do."""
    bucket = getattr(model, "_sam2b200_grad_bucket", None)
    if bucket is None or world <= 1:
        if world > 1:
            allreduce_gradients(model, world)
        return None
    if not bucket.owns_all():
        allreduce_gradients(model, world)       # re-attaches the views (see there); synchronous on the compute stream
        return None
    return bucket.allreduce(world, async_op=True)


def allreduce_gradients(model: torch.nn.Module, world: int):
    """Average gradients over ranks.  Uses the model's GradBucket if one is attached (no copies),
    else flattens on the fly."""
    bucket = getattr(model, "_sam2b200_grad_bucket", None)
    if bucket is not None:
        if not bucket.owns_all():
            # optimizer.zero_grad(set_to_none=True) (the PyTorch default) detached the views: the stack then fell back to
            # autograd, which wrote FRESH .grad tensors -- reducing the stale flat buffer would silently leave the ranks
            # out of sync.  Pull those gradients into the bucket (parameters without a gradient contribute zeros) and
            # re-install the views, so the next backward accumulates in place again.
            stale = [p for p in bucket.params if not bucket.owns(p)]
            with torch.no_grad():
                for p in stale:
                    o = bucket.offset[id(p)]
                    if p.grad is None:
                        bucket.flat[o:o + p.numel()].zero_()
            bucket.attach()
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
    """Give `model` a flat gradient buffer.  A model that offers `grad_bucket_order()` (MemoryAttention) chooses the
    layout and, from then on, accumulates its gradients into the buffer directly inside its backward kernels.

    Direct accumulation hides the parameters from autograd (they are passed detached), so torch / Lightning DDP reducer
    hooks never fire for them: use ``allreduce_gradients`` (one NCCL call on the flat buffer) instead of wrapping the
    module in ``DistributedDataParallel`` -- ``attach_grad_bucket`` refuses a DDP-wrapped module."""
    if isinstance(model, torch.nn.parallel.DistributedDataParallel):
        raise ValueError("attach_grad_bucket: direct gradient accumulation cannot be combined with a DistributedDataParallel "
                         "wrapper (its reducer hooks would never fire); attach to the bare module and call allreduce_gradients")
    order = model.grad_bucket_order() if hasattr(model, "grad_bucket_order") else None
    b = GradBucket(list(model.parameters()), order=order)
    model._sam2b200_grad_bucket = b
    return b
