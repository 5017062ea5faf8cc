"""torch.autograd.Functions over the C ABI (include/sam2_b200.h).  Device memory, streams and
autograd plumbing only -- all arithmetic of the attention core happens in libsam2b200.so."""
from __future__ import annotations

import math

import torch

from . import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}

# Optional per-kernel-family device timing (bench.py): name -> list of (start_event, end_event,
# algorithmic_flops).  Events are recorded on the launching stream; nothing synchronises here.
PROFILE = None


class _Timed:
    def __init__(self, name, flops=0.0):
        self.name, self.flops = name, flops

    def __enter__(self):
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e.record()
            PROFILE.setdefault(self.name, []).append((self.s, self.e, self.flops))
        return False


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def rope_apply(x: torch.Tensor, table: torch.Tensor, n_rope: int, inverse: bool = False,
               out_dtype=torch.bfloat16) -> torch.Tensor:
    """x: [B, L, 256] fp32|bf16 contiguous -> rotated copy (rows >= n_rope copied)."""
    if not x.is_cuda:
        raise _lib.Sam2B200Error("rope_apply: CUDA tensor required (no CPU fallback)")
    if x.dtype not in _DT:
        x = x.float()
    x = x.contiguous()
    b, l, d = x.shape
    assert d == 256 and table.dtype == torch.float32 and table.is_contiguous()
    out = torch.empty((b, l, d), dtype=out_dtype, device=x.device)
    rc = _lib.load().sam2b200_rope_apply(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _DT[out_dtype],
                                         table.data_ptr(), b, l, n_rope, table.shape[0], int(inverse),
                                         _stream(x.device))
    _lib.check(rc, "sam2b200_rope_apply")
    return out


def _drop_args(drop):
    """drop = None | (p, seed int64 device tensor [1], site) -> (p, seed pointer, site) for the C ABI."""
    if drop is None or drop[0] <= 0.0:
        return 0.0, None, 0
    p_, seed, site = drop
    assert seed.dtype == torch.int64 and seed.is_cuda and seed.numel() == 1
    return float(p_), seed.data_ptr(), int(site)


def attn_fwd(q, k, v, scale: float, nsplit: int = 0, keep_f32: bool = True, drop=None):
    """q: [B,N,256], k, v: [B,M,256] bf16 contiguous -> (out bf16 [B,N,256], out fp32 or None, lse2 fp32 [B,N]).
    drop = (p, seed tensor, site): attention-probability dropout (transformer.py:304-306)."""
    lib = _lib.load()
    b, n, d = q.shape
    m = k.shape[1]
    assert d == 256 and k.shape == v.shape and k.shape[0] == b and k.shape[2] == 256
    for t in (q, k, v):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.is_cuda
    if nsplit <= 0:
        nsplit = lib.sam2b200_attn_default_nsplit(b, n, m)
    out = torch.empty_like(q)
    out32 = torch.empty((b, n, 256), dtype=torch.float32, device=q.device) if keep_f32 else None
    lse2 = torch.empty((b, n), dtype=torch.float32, device=q.device)
    wsb = lib.sam2b200_attn_fwd_workspace_bytes(b, n, m, nsplit)
    ws = torch.empty(wsb // 4, dtype=torch.float32, device=q.device) if wsb else None
    with _Timed("attn_fwd", 4.0 * b * n * m * 256):
        rc = lib.sam2b200_attn_fwd_ex(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                      out32.data_ptr() if out32 is not None else None, lse2.data_ptr(),
                                      ws.data_ptr() if ws is not None else None, wsb, b, n, m, scale, nsplit,
                                      *_drop_args(drop), _stream(q.device))
    _lib.check(rc, "sam2b200_attn_fwd_ex")
    return out, out32, lse2


def attn_fwd_proj(q, k, v, w16, bias32, scale: float, keep_f32: bool = True, drop=None):
    """attn_fwd with the output projection fused into the kernel's epilogue (sam2b200_attn_fwd_proj; no split-KV):
    returns (out bf16, out fp32 | None, lse2, proj bf16 [B, N, 256] = out @ w16^T + bias32)."""
    lib = _lib.load()
    b, n, d = q.shape
    m = k.shape[1]
    assert d == 256 and k.shape == v.shape == (b, m, 256) and w16.shape == (256, 256) and bias32.shape == (256,)
    for t in (q, k, v, w16):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.is_cuda
    assert bias32.dtype == torch.float32 and bias32.is_contiguous()
    out = torch.empty_like(q)
    out32 = torch.empty((b, n, 256), dtype=torch.float32, device=q.device) if keep_f32 else None
    lse2 = torch.empty((b, n), dtype=torch.float32, device=q.device)
    proj = torch.empty((b, n, 256), dtype=torch.bfloat16, device=q.device)
    with _Timed("attn_fwd", 4.0 * b * n * m * 256):
        rc = lib.sam2b200_attn_fwd_proj(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                        out32.data_ptr() if out32 is not None else None, lse2.data_ptr(), w16.data_ptr(),
                                        bias32.data_ptr(), proj.data_ptr(), b, n, m, scale, *_drop_args(drop), _stream(q.device))
    _lib.check(rc, "sam2b200_attn_fwd_proj")
    return out, out32, lse2, proj


def attn_fwd_v64_proj(q, k, memv, w16, bias32, rank1_32, scale: float, keep_f32: bool = True, drop=None):
    """attn_fwd_v64 with the folded projection out_proj(v_proj(.)) fused into the epilogue (sam2b200_attn_fwd_v64_proj):
    returns (out64 bf16, out64 fp32 | None, lse2, rowsum | None, proj bf16 [B, N, 256] = out64 @ w16^T + bias32 (+ rowsum x rank1_32))."""
    lib = _lib.load()
    b, n, d = q.shape
    m = k.shape[1]
    assert d == 256 and k.shape == (b, m, 256) and memv.shape == (b, m, 64) and w16.shape == (256, 64)
    for t in (q, k, memv, w16):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.is_cuda
    assert bias32.dtype == torch.float32 and bias32.shape == (256,) and (rank1_32 is None or rank1_32.dtype == torch.float32)
    out = torch.empty((b, n, 64), dtype=torch.bfloat16, device=q.device)
    out32 = torch.empty((b, n, 64), dtype=torch.float32, device=q.device) if keep_f32 else None
    lse2 = torch.empty((b, n), dtype=torch.float32, device=q.device)
    rowsum = torch.empty((b, n), dtype=torch.float32, device=q.device) if drop is not None else None
    proj = torch.empty((b, n, 256), dtype=torch.bfloat16, device=q.device)
    with _Timed("attn_fwd", 4.0 * b * n * m * 256):
        rc = lib.sam2b200_attn_fwd_v64_proj(q.data_ptr(), k.data_ptr(), memv.data_ptr(), out.data_ptr(),
                                            out32.data_ptr() if out32 is not None else None, lse2.data_ptr(),
                                            rowsum.data_ptr() if rowsum is not None else None, w16.data_ptr(), bias32.data_ptr(),
                                            rank1_32.data_ptr() if rank1_32 is not None else None, proj.data_ptr(), b, n, m, scale,
                                            *_drop_args(drop), _stream(q.device))
    _lib.check(rc, "sam2b200_attn_fwd_v64_proj")
    return out, out32, lse2, rowsum, proj


def attn_bwd(q, k, v, out, out32, dout, lse2, scale: float, table=None, n_rope_k: int = 0, grad_dtype=torch.float32,
             dq=None, dk=None, dv=None, dbias=(None, None, None), parts: int = 0, delta=None, drop=None):
    """Backward of the attention core.  With `table` the conjugate RoPE is fused into the epilogue (dq / dk are then
    gradients w.r.t. the un-rotated projections).  dbias = (dbq, dbk, dbv): optional fp32 [256] tensors the column sums
    of dq / dk / dv are ADDED to (bias gradients of the projections; fp32 atomics).  dq / dk / dv may be preallocated 2-D/3-D views whose last dim is
    contiguous (row stride = .stride(-2)), e.g. column slices of one [R, 768] buffer."""
    lib = _lib.load()
    b, n, _ = q.shape
    m = k.shape[1]
    dev = q.device
    need = (lambda bit: parts == 0 or bool(parts & bit))
    if dq is None and need(8):
        dq = torch.empty((b, n, 256), dtype=grad_dtype, device=dev)
    if dk is None and need(4):
        dk = torch.empty((b, m, 256), dtype=grad_dtype, device=dev)
    if dv is None and need(2):
        dv = torch.empty((b, m, 256), dtype=grad_dtype, device=dev)
    for t_ in (dq, dk, dv):
        assert t_ is None or (t_.dtype == grad_dtype and t_.stride(-1) == 1)
    if delta is None:
        delta = torch.empty((b, n), dtype=torch.float32, device=dev)
    # parts (bit mask, 0 = all): 1 Delta, 2 dV, 4 dK, 8 dQ; algorithmic FLOPs: dV 2 GEMMs... counted as 4 / 4 / 2 of the 10 N M d
    work = 10.0 * b * n * m * 256 * ((4 if parts in (0,) or parts & 8 else 0) + (2 if parts in (0,) or parts & 2 else 0)
                                     + (4 if parts in (0,) or parts & 4 else 0)) / 10.0
    with _Timed("attn_bwd", work):
        for t_ in dbias:
            assert t_ is None or (t_.dtype == torch.float32 and t_.numel() == 256 and t_.is_contiguous())
        rc = lib.sam2b200_attn_bwd_ex(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr() if out is not None else None,
                                   out32.data_ptr() if out32 is not None else None, dout.data_ptr(),
                                   lse2.data_ptr(), delta.data_ptr(), *[t_.data_ptr() if t_ is not None else None for t_ in (dq, dk, dv)],
                                   _DT[grad_dtype], *[t_.stride(-2) if t_ is not None else 256 for t_ in (dq, dk, dv)],
                                   table.data_ptr() if table is not None else None,
                                   table.shape[0] if table is not None else 0, n_rope_k,
                                   b, n, m, scale, *[t_.data_ptr() if t_ is not None else None for t_ in dbias],
                                   int(parts), *_drop_args(drop), _stream(dev))
    _lib.check(rc, "sam2b200_attn_bwd_ex")
    return dq, dk, dv


def attn_fwd_v64(q, k, memv, scale: float, keep_f32: bool = True, drop=None):
    """Cross-attention on the raw 64-d memory features (sam2b200_attn_fwd_v64): q [B,N,256], k [B,M,256] (rotated),
    memv [B,M,64] bf16 -> (out64 bf16 [B,N,64], its fp32 copy or None, lse2 [B,N], rowsum); the caller applies v_proj to
    out64.  drop = (p, seed, site): rowsum [B,N] = row sums of the dropped, re-scaled probabilities (None without dropout)."""
    lib = _lib.load()
    b, n, d = q.shape
    m = k.shape[1]
    assert d == 256 and k.shape == (b, m, 256) and memv.shape == (b, m, 64)
    for t in (q, k, memv):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.is_cuda
    out = torch.empty((b, n, 64), dtype=torch.bfloat16, device=q.device)
    out32 = torch.empty((b, n, 64), dtype=torch.float32, device=q.device) if keep_f32 else None
    lse2 = torch.empty((b, n), dtype=torch.float32, device=q.device)
    rowsum = torch.empty((b, n), dtype=torch.float32, device=q.device) if drop is not None else None
    with _Timed("attn_fwd", 4.0 * b * n * m * 256):      # algorithmic FLOPs of the op it replaces
        rc = lib.sam2b200_attn_fwd_v64(q.data_ptr(), k.data_ptr(), memv.data_ptr(), out.data_ptr(),
                                       out32.data_ptr() if out32 is not None else None, lse2.data_ptr(),
                                       rowsum.data_ptr() if rowsum is not None else None, b, n, m, scale,
                                       *_drop_args(drop), _stream(q.device))
    _lib.check(rc, "sam2b200_attn_fwd_v64")
    return out, out32, lse2, rowsum


def attn_bwd_v64(q, k, memv, dout64, lse2, delta, scale: float, table=None, n_rope_k: int = 0, grad_dtype=torch.float32,
                 dq=None, dk=None, dbias=(None, None), parts: int = 12, dp_bias=None, drop=None):
    """Backward of attn_fwd_v64: dout64 = dO Wv [B,N,64] bf16, delta = rowsum(dout64 o out64) [B,N] fp32.
    parts: 4 = dK, 8 = dQ (no dV on this path).  Returns (dq, dk)."""
    lib = _lib.load()
    b, n, _ = q.shape
    m = k.shape[1]
    dev = q.device
    assert dout64.shape == (b, n, 64) and dout64.dtype == torch.bfloat16 and dout64.is_contiguous()
    assert delta.shape == (b, n) and delta.dtype == torch.float32 and delta.is_contiguous()
    if dq is None and parts & 8:
        dq = torch.empty((b, n, 256), dtype=grad_dtype, device=dev)
    if dk is None and parts & 4:
        dk = torch.empty((b, m, 256), dtype=grad_dtype, device=dev)
    for t_ in (dq, dk):
        assert t_ is None or (t_.dtype == grad_dtype and t_.stride(-1) == 1)
    work = 10.0 * b * n * m * 256 * ((4 if parts & 8 else 0) + (6 if parts & 4 else 0)) / 10.0   # dK here stands for the whole key side
    with _Timed("attn_bwd", work):
        rc = lib.sam2b200_attn_bwd_v64(q.data_ptr(), k.data_ptr(), memv.data_ptr(), dout64.data_ptr(), lse2.data_ptr(),
                                       delta.data_ptr(), dq.data_ptr() if dq is not None else None,
                                       dk.data_ptr() if dk is not None else None, _DT[grad_dtype],
                                       dq.stride(-2) if dq is not None else 256, dk.stride(-2) if dk is not None else 256,
                                       table.data_ptr() if table is not None else None,
                                       table.shape[0] if table is not None else 0, n_rope_k, b, n, m, scale,
                                       *[t_.data_ptr() if t_ is not None else None for t_ in dbias], int(parts),
                                       dp_bias.data_ptr() if dp_bias is not None else None, *_drop_args(drop), _stream(dev))
    _lib.check(rc, "sam2b200_attn_bwd_v64")
    return dq, dk


class RopeAttentionFn(torch.autograd.Function):
    """out = softmax(rope(q) rope(k[:, :M-P])^T / sqrt(256)) v for one 256-wide head.

    Replaces transformer.py:296-306 of the reference (apply_rotary_enc + slice write-back + SDPA)
    and its autograd backward.  q: [B, N, 256]; k, v: [B, M, 256]; bf16 out."""

    @staticmethod
    def forward(ctx, q, k, v, table, num_k_exclude_rope: int, nsplit: int, drop_p: float = 0.0):
        n_rope_k = k.shape[1] - num_k_exclude_rope
        # attention-probability dropout (train mode, transformer.py:304-306): a fresh device seed per call
        ctx.drop = None
        if drop_p > 0.0:
            ctx.drop = (float(drop_p), torch.empty(1, dtype=torch.int64, device=q.device).random_(), 0)
        scale = 1.0 / math.sqrt(q.shape[-1])
        q_rot = rope_apply(q, table, q.shape[1])
        k_rot = rope_apply(k, table, n_rope_k)
        vb = v.to(torch.bfloat16).contiguous()
        out, out32, lse2 = attn_fwd(q_rot, k_rot, vb, scale, nsplit, drop=ctx.drop)
        ctx.save_for_backward(q_rot, k_rot, vb, out32, lse2, table)
        ctx.scale = scale
        ctx.n_rope_k = n_rope_k
        ctx.in_dtypes = (q.dtype, k.dtype, v.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        q_rot, k_rot, vb, out32, lse2, table = ctx.saved_tensors
        dout = dout.to(torch.bfloat16).contiguous()
        gdt = torch.bfloat16 if all(d == torch.bfloat16 for d in ctx.in_dtypes) else torch.float32
        dq, dk, dv = attn_bwd(q_rot, k_rot, vb, None, out32, dout, lse2, ctx.scale, table=table,
                              n_rope_k=ctx.n_rope_k, grad_dtype=gdt, drop=ctx.drop)
        return dq.to(ctx.in_dtypes[0]), dk.to(ctx.in_dtypes[1]), dv.to(ctx.in_dtypes[2]), None, None, None, None


def _out_dt(dt):
    return dt if dt in _DT else torch.float32
