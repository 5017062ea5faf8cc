"""CUDA-graph replay of the MemoryAttention forward + backward.

The per-frame call launches ~350 small-to-medium kernels (4 layers x [LayerNorm + projections, attention, MLP] forward
and backward); at 384 px the host-side launch cost equals the GPU time.  The memory bank takes only a handful of shapes
during a clip (M = min(t, 7) N + 4 min(t, 16)), so each distinct signature is captured ONCE -- a forward graph and a
backward graph sharing one private memory pool -- and replayed afterwards: the host submits two graph launches per frame
instead of hundreds of kernels.  SURVEY.md section 8f rank 4 ("CUDA-graph the per-frame step").

The capture calls ``MemoryAttentionStackFn.forward`` / ``.backward`` DIRECTLY with a stand-in context object: no autograd
engine runs inside a capture, so no AccumulateGrad node is ever bound to a capture stream (round 1 went through
``torch.cuda.make_graphed_callables``, whose warm-up and capture streams differ: PyTorch warned about the stream mismatch
and a model that had already run an eager backward could not be captured at all).

Usage:  fast = GraphedMemoryAttention(memory_attention)   # same call signature as MemoryAttention
"""
from __future__ import annotations

import os

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import fused_stack
from . import ops as _ops


class _Ctx:
    """What MemoryAttentionStackFn.forward / backward need from an autograd context."""

    def __init__(self, needs_input_grad):
        self.needs_input_grad = tuple(needs_input_grad)
        self.saved_tensors = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tuple(tensors)


class _Captured:
    """One (forward graph, backward graph) pair with its static input / output tensors."""

    def __init__(self, module, tensors, p, bank_meta, n_extra):
        inner = module
        dev = tensors[0].device
        self.static_in = [t.detach().clone() for t in tensors]              # curr, memory, curr_pos, memory_pos, [tpos, objpos]
        self.req = [bool(t.requires_grad) for t in tensors]
        self.params = [q for _, q in inner.named_parameters()]
        bucket = getattr(inner, "_sam2b200_grad_bucket", None)
        self.direct = fused_stack.direct_grads_possible(bucket, self.params)
        self.anchor_used = self.direct
        s_curr, s_mem, s_cpos, s_mpos = self.static_in[:4]
        extra = self.static_in[4:4 + n_extra]

        def build():
            packed = (extra[0], extra[1], bank_meta[0], bank_meta[1]) if bank_meta is not None else None
            return inner._forward_fused(s_curr, s_mem, s_cpos, s_mpos, int(p), packed=packed, raw=True, direct=self.direct,
                                        anchor=torch.zeros(1, device=dev) if self.direct else None)

        def needs(args):
            # inputs of the function: (meta, curr, curr_pos, memory, memory_pos, bank_tpos, bank_objpos, *params[, anchor])
            r = dict(zip(("curr", "mem", "cpos", "mpos"), self.req[:4]))
            ex = self.req[4:4 + n_extra] + [False, False]
            flags = [False, r["curr"], r["cpos"], r["mem"], r["mpos"], ex[0] and args[5] is not None, ex[1] and args[6] is not None]
            n_par = len(self.params)
            flags += [(not self.direct) and q.requires_grad for q in self.params]
            if self.direct:
                flags.append(True)
            assert len(flags) == len(args) and n_par
            return flags

        fn = fused_stack.MemoryAttentionStackFn
        saved_profile, _ops.PROFILE = _ops.PROFILE, None          # events cannot be timed inside a capture
        fused_stack.bf16_params(self.params)                        # mirror refreshed OUTSIDE the capture
        kept = bucket.flat.clone() if bucket is not None else None  # the warm-up / capture backward accumulates into the bucket
        try:
            torch.cuda.synchronize()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():          # warm-up: lazy initialisation must not land in a capture
                for _ in range(2):
                    args = build()
                    ctx = _Ctx(needs(args))
                    out = fn.forward(ctx, *args)
                    fn.backward(ctx, torch.zeros_like(out))
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            pool = torch.cuda.graph_pool_handle()
            self.fwd, self.bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            # Experiment (SAM2B200_GRAPH_PRIORITY=1, NOT faster: 40.61 / 40.83 vs 40.49 / 40.70 ms per cfg2 step): capture on a high-priority
            # stream, so that the kernel nodes of the critical path outrank the side stream's weight-gradient work
            cap = torch.cuda.Stream(device=dev, priority=-1) if os.environ.get("SAM2B200_GRAPH_PRIORITY") else None
            kw = dict(pool=pool) if cap is None else dict(pool=pool, stream=cap)
            with torch.no_grad():
                with torch.cuda.graph(self.fwd, **kw):
                    args = build()
                    self.ctx = _Ctx(needs(args))
                    self.static_out = fn.forward(self.ctx, *args)
                self.static_gout = torch.zeros_like(self.static_out)
                with torch.cuda.graph(self.bwd, **kw):
                    self.static_grads = fn.backward(self.ctx, self.static_gout)
        finally:
            _ops.PROFILE = saved_profile
            if kept is not None:
                torch.cuda.synchronize()
                bucket.flat.copy_(kept)
        self.n_extra = n_extra


class _Replay(torch.autograd.Function):
    """forward(cap, curr, memory, curr_pos, memory_pos, [tpos, objpos], *params) -> out (a static tensor of the graph)."""

    @staticmethod
    def forward(ctx, cap: _Captured, *inputs):
        n_in = 4 + cap.n_extra
        for s, t in zip(cap.static_in, inputs[:n_in]):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t)
        cap.fwd.replay()
        ctx.cap = cap
        ctx.n_in = n_in
        return cap.static_out.detach()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        cap = ctx.cap
        if cap.static_gout.data_ptr() != gout.data_ptr():
            cap.static_gout.copy_(gout)
        cap.bwd.replay()
        g = cap.static_grads           # (None, d_curr, d_pos, d_mem, d_mpos, d_tpos, d_objpos, *param grads[, anchor])
        d_curr, d_cpos, d_mem, d_mpos, d_tpos, d_obj = g[1], g[2], g[3], g[4], g[5], g[6]
        outs = [None, d_curr, d_mem, d_cpos, d_mpos] + [d_tpos, d_obj][:cap.n_extra]
        if cap.direct:
            outs += [None] * len(cap.params)          # accumulated straight into the GradBucket by the captured kernels
        else:
            outs += list(g[7:7 + len(cap.params)])
        return tuple(outs)


class GraphedMemoryAttention(nn.Module):
    def __init__(self, inner: nn.Module, max_signatures: int = 32):
        super().__init__()
        self.inner = inner
        self.max_signatures = max_signatures
        self._graphs: Dict[Tuple, _Captured] = {}

    @property
    def layers(self):
        return self.inner.layers

    def _signature(self, curr, memory, curr_pos, memory_pos, p, extra=(), bank_meta=None) -> Tuple:
        params = [q for _, q in self.inner.named_parameters()]
        direct = fused_stack.direct_grads_possible(getattr(self.inner, "_sam2b200_grad_bucket", None), params)
        l0 = self.inner.layers[0]
        return (direct, bank_meta, tuple(curr.shape), tuple(memory.shape), curr.dtype, memory.dtype, curr.requires_grad,
                curr_pos.requires_grad, memory.requires_grad, memory_pos.requires_grad, int(p), self.inner.training,
                (l0.dropout_value, l0.self_attn.dropout_p, l0.cross_attn_image.dropout_p),
                tuple(q.requires_grad for q in params), tuple((tuple(e.shape), e.requires_grad) for e in extra))

    def forward(self, curr, memory, curr_pos: Optional[Tensor] = None, memory_pos: Optional[Tensor] = None,
                num_obj_ptr_tokens: int = 0):
        if isinstance(curr, list):
            assert isinstance(curr_pos, list) and len(curr) == len(curr_pos) == 1
            curr, curr_pos = curr[0], curr_pos[0]
        from .memory_bank import PackedBank
        bank = memory if isinstance(memory, PackedBank) else None
        extra, bank_meta = (), None
        if bank is not None:
            # the graphs take the two packed tensors where the reference layout has memory / memory_pos, plus the two small
            # differentiable position tensors of the bank (an empty tensor stands for an absent one)
            memory, memory_pos, num_obj_ptr_tokens = bank.memk, bank.memv, bank.num_obj_ptr_tokens
            empty = lambda: torch.zeros((0, 64), dtype=torch.float32, device=curr.device)
            extra = (bank.tpos_rows if bank.tpos_rows is not None else empty(), bank.obj_pos if bank.obj_pos is not None else empty())
            bank_meta = (int(bank.n_slots), int(bank.hw))
        eligible = (curr.is_cuda and curr_pos is not None and memory_pos is not None and torch.is_grad_enabled()
                    and not torch.cuda.is_current_stream_capturing()
                    and getattr(self.inner, "use_fused_stack", False) and self.inner._fused_eligible())
        if not eligible:
            return self.inner(curr, bank if bank is not None else memory, curr_pos, memory_pos, num_obj_ptr_tokens)
        key = self._signature(curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens, extra, bank_meta)
        cap = self._graphs.get(key)
        if cap is None:
            if len(self._graphs) >= self.max_signatures:
                return self.inner(curr, bank if bank is not None else memory, curr_pos, memory_pos, num_obj_ptr_tokens)
            cap = _Captured(self.inner, (curr, memory, curr_pos, memory_pos, *extra), num_obj_ptr_tokens, bank_meta, len(extra))
            self._graphs[key] = cap
        # the graphs read the persistent bf16 weight mirror: bring it up to date (no-op unless a parameter changed)
        fused_stack.bf16_params(cap.params)
        tail = list(cap.params)
        if cap.direct:
            # the parameters are hidden from autograd (their gradients go straight into the GradBucket): an anchor leaf that
            # requires grad makes sure the backward runs even when no input asks for a gradient
            tail = [self.inner._grad_anchor(curr.device)]
            return _ReplayDirect.apply(cap, curr, memory, curr_pos, memory_pos, *extra, *tail)
        return _Replay.apply(cap, curr, memory, curr_pos, memory_pos, *extra, *tail)


class _ReplayDirect(torch.autograd.Function):
    """Direct-gradient mode: inputs (cap, curr, memory, curr_pos, memory_pos, [tpos, objpos], anchor)."""

    @staticmethod
    def forward(ctx, cap: _Captured, *inputs):
        return _Replay.forward(ctx, cap, *inputs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        cap = ctx.cap
        if cap.static_gout.data_ptr() != gout.data_ptr():
            cap.static_gout.copy_(gout)
        cap.bwd.replay()
        g = cap.static_grads
        return (None, g[1], g[3], g[2], g[4], *[g[5], g[6]][:cap.n_extra], None)
