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
        return (tuple(curr.shape), tuple(memory.shape), curr.dtype, memory.dtype, curr.requires_grad,
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
        return g(curr, memory, curr_pos, memory_pos)

    def _capture(self, curr, memory, curr_pos, memory_pos, p):
        mod = _FixedPointerCount(self.inner, p)
        sample = tuple(torch.randn_like(t).requires_grad_(t.requires_grad) for t in (curr, memory, curr_pos, memory_pos))
        saved_profile = _ops.PROFILE
        _ops.PROFILE = None                      # events cannot be timed inside a capture
        fused_stack.CAPTURE_SAFE_CASTS = True    # weight casts are recorded into the graph, never cached
        try:
            graphed = torch.cuda.make_graphed_callables(mod, sample, num_warmup_iters=2, allow_unused_input=True)
        finally:
            fused_stack.CAPTURE_SAFE_CASTS = False
            _ops.PROFILE = saved_profile
        return graphed
