"""CUDA-graph replay of the MemoryAttention forward + backward.

The per-frame call launches ~500 small-to-medium kernels (4 layers x [LayerNorm, projections, RoPE,
attention, MLP] forward and backward); at 384 px the host-side launch cost equals the GPU time.  The
memory bank takes only a handful of shapes during a clip (M = min(t,7)*(N+4)), so each distinct
signature is captured ONCE -- forward graph and backward graph sharing one private memory pool
(torch.cuda.make_graphed_callables) -- and replayed afterwards: the host submits two graph launches
per frame instead of hundreds of kernels.  SURVEY.md section 8f rank 4 ("CUDA-graph the per-frame step").

Usage:  fast = GraphedMemoryAttention(memory_attention)   # same call signature as MemoryAttention
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import fused_stack
from . import ops as _ops


class _FixedPointerCount(nn.Module):
    """MemoryAttention with num_obj_ptr_tokens bound (graph capture needs tensor-only arguments)."""

    def __init__(self, inner: nn.Module, num_obj_ptr_tokens: int):
        super().__init__()
        self.inner = inner
        self.p = int(num_obj_ptr_tokens)

    def forward(self, curr, memory, curr_pos, memory_pos):
        return self.inner(curr, memory, curr_pos, memory_pos, self.p)


class GraphedMemoryAttention(nn.Module):
    def __init__(self, inner: nn.Module, max_signatures: int = 32):
        super().__init__()
        self.inner = inner
        self.max_signatures = max_signatures
        self._graphs: Dict[Tuple, object] = {}

    @property
    def layers(self):
        return self.inner.layers

    def _signature(self, curr, memory, curr_pos, memory_pos, p) -> Tuple:
        direct = fused_stack.direct_grads_possible(getattr(self.inner, "_sam2b200_grad_bucket", None),
                                                   [q for _, q in self.inner.named_parameters()])
        return (direct, tuple(curr.shape), tuple(memory.shape), curr.dtype, memory.dtype, curr.requires_grad,
                curr_pos.requires_grad, memory.requires_grad, memory_pos.requires_grad, int(p), self.inner.training,
                torch.is_grad_enabled())

    def forward(self, curr, memory, curr_pos: Optional[Tensor] = None, memory_pos: Optional[Tensor] = None,
                num_obj_ptr_tokens: int = 0):
        if isinstance(curr, list):
            assert isinstance(curr_pos, list) and len(curr) == len(curr_pos) == 1
            curr, curr_pos = curr[0], curr_pos[0]
        eligible = (curr.is_cuda and curr_pos is not None and memory_pos is not None and torch.is_grad_enabled()
                    and getattr(self.inner, "use_fused_stack", False) and self.inner._fused_eligible())
        if not eligible:
            return self.inner(curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens)
        key = self._signature(curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self.max_signatures:
                return self.inner(curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens)
            g = self._capture(curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens)
            self._graphs[key] = g
        # the graphs read the persistent bf16 weight mirror: bring it up to date (no-op unless a parameter changed)
        fused_stack.bf16_params([p for _, p in self.inner.named_parameters()])
        return g(curr, memory, curr_pos, memory_pos)

    def _capture(self, curr, memory, curr_pos, memory_pos, p):
        params = [q for _, q in self.inner.named_parameters()]
        bucket = getattr(self.inner, "_sam2b200_grad_bucket", None)
        if fused_stack.direct_grads_possible(bucket, params):
            # gradients go straight into the GradBucket: the parameters are not part of the graphed input surface
            # (a plain function, not an nn.Module), so autograd never touches them at capture or replay time
            inner = self.inner
            direct = True

            def mod(curr, memory, curr_pos, memory_pos, anchor):
                return inner._forward_fused(curr, memory, curr_pos, memory_pos, int(p), anchor=anchor)
        else:
            direct = False
            mod = _FixedPointerCount(self.inner, p)
        sample = tuple(torch.randn_like(t).requires_grad_(t.requires_grad) for t in (curr, memory, curr_pos, memory_pos))
        if direct:
            sample = sample + (torch.zeros(1, device=curr.device, requires_grad=True),)
        saved_profile = _ops.PROFILE
        _ops.PROFILE = None                      # events cannot be timed inside a capture
        fused_stack.bf16_params([p for _, p in self.inner.named_parameters()])   # mirror refreshed OUTSIDE the capture
        # warm-up and capture run the backward on random inputs; with a GradBucket attached that backward accumulates
        # straight into the bucket, so its content is put back afterwards
        kept = bucket.flat.clone() if bucket is not None else None
        try:
            graphed = torch.cuda.make_graphed_callables(mod, sample, num_warmup_iters=2, allow_unused_input=True)
        finally:
            _ops.PROFILE = saved_profile
            if kept is not None:
                torch.cuda.synchronize()
                bucket.flat.copy_(kept)
        if direct:
            anchor = self.inner._grad_anchor(curr.device)
            return lambda c, m, cp, mp: graphed(c, m, cp, mp, anchor)
        return graphed
