"""Whole-stack autograd.Function for SAM2 MemoryAttention on B200.

One hand-scheduled forward / backward for the 4-layer stack of
sam2_video/model/modeling/memory_attention.py:119-169 in its shipped configuration
(configs/sam2/sam2.1_hiera_t.yaml:29-60).  Compared with composing nn.Modules under autograd it
  * keeps the residual stream in fp32 and fuses residual-add + LayerNorm + bf16 cast (ln_fwd),
    LayerNorm backward + residual-gradient add (ln_bwd), cast + bias-gradient (colsum) into single
    passes of the hand-written kernels in csrc/glue.cu,
  * runs the attention forward (with the output projection in its epilogue) and backward in the tcgen05 kernels of
    csrc/attn*.cuh -- the cross-attention on the raw 64-d memory features,
  * runs every dense GEMM but the two MLP weight gradients on our own tcgen05 kernels with the neighbouring element-wise work in
    their epilogues: sam2b200_gemm / gemm_ex (csrc/gemm.cu: projections with bias, RoPE, ReLU + dropout; input gradients; the
    Delta row sums), sam2b200_wgrad / wgrad2 (csrc/wgrad.cu: weight gradients accumulated in place, bias gradients, bank-segment
    sums), sam2b200_mlp_dh (csrc/mlp.cu),
  * packs `memory + pos` once per call instead of once per layer (memory_attention.py:75-76), or takes the bank already packed
    (memory_bank.PackedBank),
  * puts everything that only feeds parameter gradients on a second stream.
PyTorch is used for device memory, streams and the few remaining library calls.
"""
from __future__ import annotations

import math
import os
from typing import List

import torch

from . import _lib
from .ops import _Timed, attn_bwd, attn_bwd_v64, attn_fwd, attn_fwd_proj, attn_fwd_v64, attn_fwd_v64_proj, rope_apply

BF16 = torch.bfloat16
F32 = torch.float32

# per-layer parameter order == reference named_parameters() order (26 tensors per layer)
_LAYER_KEYS = [
    "sa.q.w", "sa.q.b", "sa.k.w", "sa.k.b", "sa.v.w", "sa.v.b", "sa.o.w", "sa.o.b",
    "ca.q.w", "ca.q.b", "ca.k.w", "ca.k.b", "ca.v.w", "ca.v.b", "ca.o.w", "ca.o.b",
    "l1.w", "l1.b", "l2.w", "l2.b", "n1.w", "n1.b", "n2.w", "n2.b", "n3.w", "n3.b",
]
_NPL = len(_LAYER_KEYS)


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _drop(drop):
    """drop = None | (p, seed int64 device tensor [1], site) -> (p, seed pointer, site) of the C ABI."""
    if drop is None or drop[0] <= 0.0:
        return 0.0, None, 0
    return float(drop[0]), drop[1].data_ptr(), int(drop[2])


def ln_fwd(x, res, gamma, beta, want_f32_seq_first=None, eps=1e-5, drop=None):
    """x: [R,256] fp32; res: [R,256] bf16 or None.  Returns (y, x_new, mean, rstd); y is bf16 [R,256] or,
    if want_f32_seq_first=(B, N), fp32 [N, B, 256].  drop = (p, seed, site): x_new = x + dropout(res)."""
    lib = _lib.load()
    rows = x.shape[0]
    dev = x.device
    x_new = torch.empty_like(x) if res is not None else x
    mean = torch.empty(rows, dtype=F32, device=dev)
    rstd = torch.empty(rows, dtype=F32, device=dev)
    if want_f32_seq_first is None:
        y = torch.empty((rows, 256), dtype=BF16, device=dev)
        y16, y32, tb, tn = y.data_ptr(), None, 0, 0
    else:
        tb, tn = want_f32_seq_first
        y = torch.empty((tn, tb, 256), dtype=F32, device=dev)
        y16, y32 = None, y.data_ptr()
    rc = lib.sam2b200_ln_fwd(x.data_ptr(), res.data_ptr() if res is not None else None,
                             x_new.data_ptr() if res is not None else None, gamma.data_ptr(), beta.data_ptr(),
                             y16, y32, mean.data_ptr(), rstd.data_ptr(), rows, eps, tb, tn, *_drop(drop), _stream(dev))
    _lib.check(rc, "sam2b200_ln_fwd")
    return y, x_new, mean, rstd


def ln_bwd(dy, x, mean, rstd, gamma, g_in, dgamma, dbeta, seq_first=None, dbias=None, drop=None, side=None):
    """Returns g_out (fp32) or, with dbias, (g_out, bf16 copy of g_out) and dbias += column sums of the copy.
    drop = (p, seed, site): the copy is the gradient of a branch that went through dropout -> masked and scaled.
    side (a _SideStream): the fold of the per-block partial sums into dgamma / dbeta / dbias -- parameter gradients only -- is enqueued
    on the side stream instead of sitting (24 CTAs, ~7 us) on the critical path of the residual-stream gradient."""
    lib = _lib.load()
    rows = x.shape[0]
    dev = x.device
    g_out = torch.empty_like(x)
    g16 = torch.empty(x.shape, dtype=BF16, device=dev) if dbias is not None else None
    ws = torch.empty(lib.sam2b200_ln_bwd_workspace_bytes(rows) // 4, dtype=F32, device=dev)
    tb, tn = seq_first if seq_first is not None else (0, 0)
    is16 = dy.dtype == BF16
    def call(stages):
        rc = lib.sam2b200_ln_bwd_stages(dy.data_ptr() if is16 else None, None if is16 else dy.data_ptr(), x.data_ptr(),
                                        mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                        g_in.data_ptr() if g_in is not None else None, g_out.data_ptr(),
                                        g16.data_ptr() if g16 is not None else None, dgamma.data_ptr(), dbeta.data_ptr(),
                                        dbias.data_ptr() if dbias is not None else None, ws.data_ptr(), rows, tb, tn, *_drop(drop),
                                        stages, _stream(dev))
        _lib.check(rc, "sam2b200_ln_bwd")
    if side is not None and side.enabled and not NO_LN_FOLD_ON_SIDE:
        call(1)
        side.run(lambda: call(2), ws, dgamma, dbeta, dbias)
    else:
        call(3)
    return g_out if dbias is None else (g_out, g16)


def _colsum(mode, in32, io16, h16, colsum, rows, c, ld=0, scale=1.0):
    lib = _lib.load()
    dev = io16.device
    ws = torch.empty(max(lib.sam2b200_colsum_workspace_bytes(rows, c) // 4, 1), dtype=F32, device=dev)
    rc = lib.sam2b200_colsum(mode, in32.data_ptr() if in32 is not None else None, io16.data_ptr(),
                             h16.data_ptr() if h16 is not None else None, colsum.data_ptr(), ws.data_ptr(), rows, c,
                             ld, float(scale), _stream(dev))
    _lib.check(rc, "sam2b200_colsum")


def cast_colsum(g32, colsum):
    """bf16 copy of g32 [R,C] + colsum += column sums (bias gradient of the GEMM that consumes it)."""
    out = torch.empty(g32.shape, dtype=BF16, device=g32.device)
    _colsum(0, g32, out, None, colsum, g32.shape[0], g32.shape[1])
    return out


def relu_bwd_colsum_(dh16, h16, colsum, scale=1.0):
    """dh *= (h > 0) * scale in place; h is the (possibly dropped) activation, scale = 1 / (1 - p) of that dropout."""
    _colsum(1, None, dh16, h16, colsum, dh16.shape[0], dh16.shape[1], scale=scale)


def proj_rope(x16, w16, bias16, n_out, table=None, rope_outs=0, rows_per_item=1, n_rope_rows=0):
    """Y = x @ w^T + bias with the axial rotation of the first `rope_outs` outputs fused into the GEMM epilogue
    (sam2b200_proj_rope).  x16 [R, K] (K = 256 | 64), w16 [256 * n_out, K], bias16 [256 * n_out]; returns n_out
    contiguous [R, 256] bf16 tensors (q | k | v)."""
    r, k = x16.shape
    assert x16.is_contiguous() and w16.is_contiguous() and bias16.is_contiguous() and w16.shape == (256 * n_out, k)
    outs = [torch.empty((r, 256), dtype=BF16, device=x16.device) for _ in range(n_out)]
    ptr = [o.data_ptr() for o in outs] + [None] * (3 - n_out)
    rc = _lib.load().sam2b200_proj_rope(x16.data_ptr(), w16.data_ptr(), bias16.data_ptr(), ptr[0], ptr[1], ptr[2], r, k, 256 * n_out,
                                        256 * rope_outs, table.data_ptr() if table is not None else None, int(rows_per_item),
                                        int(n_rope_rows), table.shape[0] if table is not None else 1, _stream(x16.device))
    _lib.check(rc, "sam2b200_proj_rope")
    return outs


def ln_proj(x, res, gamma, beta, w16, bias16, n_out, out_width=256, table=None, rope_outs=0, rows_per_item=1, n_rope_rows=0,
            relu=False, drop_res=None, drop_out=None, eps=1e-5):
    """Head of a pre-norm block in ONE kernel (sam2b200_ln_proj, csrc/lnproj.cu):
        x_new = x + dropout(res);  y = LayerNorm(x_new);  outs = epi(y @ w16^T + bias16)
    x [R, 256] fp32, res [R, 256] bf16 | None, w16 [n_out * out_width, 256] bf16.  The first `rope_outs` outputs (of width
    256) are rotated with the axial `table`; relu: max(., 0) then drop_out.  Returns (outs, y, x_new, mean, rstd) -- the
    same tensors ln_fwd + addmm (+ rope_apply | dropout_inplace_) produce."""
    r = x.shape[0]
    dev = x.device
    nout = n_out * out_width
    assert x.is_contiguous() and x.dtype == F32 and w16.is_contiguous() and w16.shape == (nout, 256) and w16.dtype == BF16
    assert res is None or (res.is_contiguous() and res.dtype == BF16 and res.shape == x.shape)
    assert bias16 is None or (bias16.is_contiguous() and bias16.numel() == nout)
    x_new = torch.empty_like(x) if res is not None else x
    y = torch.empty((r, 256), dtype=BF16, device=dev)
    mean = torch.empty(r, dtype=F32, device=dev)
    rstd = torch.empty(r, dtype=F32, device=dev)
    outs = [torch.empty((r, out_width), dtype=BF16, device=dev) for _ in range(n_out)]
    ptr = [o.data_ptr() for o in outs] + [None] * (3 - n_out)
    rc = _lib.load().sam2b200_ln_proj(
        x.data_ptr(), res.data_ptr() if res is not None else None, x_new.data_ptr() if res is not None else None, gamma.data_ptr(),
        beta.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), r, float(eps), w16.data_ptr(),
        bias16.data_ptr() if bias16 is not None else None, nout, ptr[0], ptr[1], ptr[2], int(out_width), 256 * int(rope_outs),
        table.data_ptr() if table is not None else None, int(rows_per_item), int(n_rope_rows), table.shape[0] if table is not None else 1,
        int(bool(relu)), *_drop(drop_res), *_drop(drop_out), _stream(dev))
    _lib.check(rc, "sam2b200_ln_proj")
    return outs, y, x_new, mean, rstd


def gemm_ex(a16, w16, n_out, out_width, nn=False, bias=None, table=None, rope_outs=0, rows_per_item=1, n_rope_rows=0, relu=False, drop_out=None):
    """n_out bf16 tensors [R, out_width] = epi(a16 @ w16^T + bias) on sam2b200_gemm_ex (csrc/gemm.cu): w16 [n_out * out_width, K] (or
    [K, n_out * out_width] with nn=True), bias fp32; the first rope_outs outputs (width 256) rotated with the axial `table`; relu:
    max(., 0) then drop_out = (p, seed, site)."""
    r, k = a16.shape
    nout = n_out * out_width
    assert a16.dtype == BF16 and w16.dtype == BF16 and a16.stride(1) == 1 and w16.stride(1) == 1
    assert w16.shape == ((k, nout) if nn else (nout, k)), (tuple(a16.shape), tuple(w16.shape), nn)
    assert bias is None or (bias.dtype == F32 and bias.shape == (nout,) and bias.is_contiguous())
    outs = [torch.empty((r, out_width), dtype=BF16, device=a16.device) for _ in range(n_out)]
    ptr = [o.data_ptr() for o in outs] + [None] * (3 - n_out)
    rc = _lib.load().sam2b200_gemm_ex(ptr[0], ptr[1], ptr[2], int(out_width), int(out_width), a16.data_ptr(), a16.stride(0), w16.data_ptr(),
                                      w16.stride(0), int(bool(nn)), r, k, nout, bias.data_ptr() if bias is not None else None,
                                      256 * int(rope_outs), table.data_ptr() if table is not None else None, int(rows_per_item),
                                      int(n_rope_rows), table.shape[0] if table is not None else 1, int(bool(relu)), *_drop(drop_out),
                                      None, None, _stream(a16.device))
    _lib.check(rc, "sam2b200_gemm_ex")
    return outs


def ln_then_proj(x, res, gamma, beta, w16, bias32, n_out, out_width=256, table=None, rope_outs=0, rows_per_item=1, n_rope_rows=0,
                 relu=False, drop_res=None, drop_out=None, eps=1e-5):
    """Head of a pre-norm block as TWO kernels: the HBM-bound LayerNorm pass (ln_fwd: x + dropout(res), LayerNorm, bf16 y, statistics)
    and the resident-CTA GEMM with the bias / RoPE / ReLU (+ dropout) epilogue (gemm_ex).  Same results and return value as ln_proj;
    measured faster than the one-kernel version, whose one-tile-per-CTA LayerNorm prologue and 64-column chunk loop leave the HBM
    pipe half idle (profiles/r2_lnproj_vs_split.txt)."""
    y, x_new, mean, rstd = ln_fwd(x, res, gamma, beta, eps=eps, drop=drop_res)
    outs = gemm_ex(y, w16, n_out, out_width, bias=bias32, table=table, rope_outs=rope_outs, rows_per_item=rows_per_item,
                   n_rope_rows=n_rope_rows, relu=relu, drop_out=drop_out)
    return outs, y, x_new, mean, rstd


def block_head(x, res, gamma, beta, w16, bias16, bias32, n_out, **kw):
    """Pre-norm block head: ln_fwd + gemm_ex (default) or the one-kernel ln_proj (SAM2B200_LNPROJ_FUSED=1 / SAM2B200_NO_GEMM=1)."""
    if LNPROJ_FUSED or NO_GEMM:
        return ln_proj(x, res, gamma, beta, w16, bias16, n_out, **kw)
    return ln_then_proj(x, res, gamma, beta, w16, bias32, n_out, **kw)


def wgrad_(c32, a16, b16, dbias=None):
    """c32 [Mo, No] fp32 += a16[R, Mo]^T @ b16[R, No] in place (sam2b200_wgrad: split over R, tcgen05, partial tiles added with
    fp32 reductions).  Row strides may exceed the widths (column slices of wider buffers).  dbias (fp32 [Mo], optional) +=
    column sums of a16 (the bias gradient of the same layer) from the same MMAs (No = 64 or Mo <= 768 only)."""
    r, mo = a16.shape
    no = b16.shape[1]
    assert c32.shape == (mo, no) and c32.dtype == F32 and c32.stride(1) == 1
    assert a16.dtype == BF16 and b16.dtype == BF16 and a16.stride(1) == 1 and b16.stride(1) == 1 and b16.shape[0] == r
    assert dbias is None or (dbias.dtype == F32 and dbias.shape == (mo,) and dbias.is_contiguous())
    rc = _lib.load().sam2b200_wgrad(c32.data_ptr(), c32.stride(0), a16.data_ptr(), a16.stride(0), b16.data_ptr(), b16.stride(0), r, mo, no,
                                    dbias.data_ptr() if dbias is not None else None, _stream(c32.device))
    _lib.check(rc, "sam2b200_wgrad")
    return c32


def gemm(a16, w16, nn=False, bias=None, table=None, rows_per_item=1, n_rope_rows=0, dot_rows=None):
    """Dense GEMM on our own resident-CTA tcgen05 kernel (sam2b200_gemm, csrc/gemm.cu), bf16 in / bf16 out, fp32 accumulation:
        nn=False: a16 [R, K] @ w16[No, K]^T + bias   (nn.Linear forward; `table`: axial rotation of rows whose position inside
                  their item of `rows_per_item` rows is < n_rope_rows, fused on the fp32 accumulator)
        nn=True:  a16 [R, K] @ w16[K, No]            (nn.Linear input gradient)
    No = 256 or 64, K a multiple of 64; bias fp32 [No]; row strides may exceed the widths.  dot_rows (fp32 [R, 64], No = 64): also
    returns dot[r] = sum_c out[r, c] * dot_rows[r, c] (fp32 [R]) from the epilogue."""
    r, k = a16.shape
    no = w16.shape[1] if nn else w16.shape[0]
    assert a16.dtype == BF16 and w16.dtype == BF16 and a16.stride(1) == 1 and w16.stride(1) == 1
    assert w16.shape == ((k, no) if nn else (no, k)), (tuple(a16.shape), tuple(w16.shape), nn)
    assert bias is None or (bias.dtype == F32 and bias.shape == (no,) and bias.is_contiguous())
    assert table is None or (table.dtype == F32 and table.is_contiguous())
    out = torch.empty((r, no), dtype=BF16, device=a16.device)
    dot = None
    if dot_rows is not None:
        assert no == 64 and dot_rows.dtype == F32 and dot_rows.is_contiguous() and dot_rows.numel() == r * 64
        dot = torch.empty(r, dtype=F32, device=a16.device)
    rc = _lib.load().sam2b200_gemm(out.data_ptr(), no, a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), int(bool(nn)), r, k, no,
                                   bias.data_ptr() if bias is not None else None, table.data_ptr() if table is not None else None,
                                   int(rows_per_item), int(n_rope_rows), table.shape[0] if table is not None else 1,
                                   dot_rows.data_ptr() if dot_rows is not None else None, dot.data_ptr() if dot is not None else None,
                                   _stream(a16.device))
    _lib.check(rc, "sam2b200_gemm")
    return out if dot is None else (out, dot)


def wgrad2_(c32, c2_32, a16, b16, b2_16, dbias=None):
    """c32 [256, 64] += a16^T @ b16 and c2_32 [256, 64] += a16^T @ b2_16 in ONE pass over a16 [R, 256] (sam2b200_wgrad2); dbias += colsum(a16)."""
    r = a16.shape[0]
    assert a16.shape == (r, 256) and b16.shape == (r, 64) and b2_16.shape == (r, 64) and c32.shape == (256, 64) and c2_32.shape == (256, 64)
    assert all(t.dtype == BF16 and t.stride(1) == 1 for t in (a16, b16, b2_16)) and all(t.dtype == F32 and t.stride(1) == 1 for t in (c32, c2_32))
    rc = _lib.load().sam2b200_wgrad2(c32.data_ptr(), c32.stride(0), c2_32.data_ptr(), c2_32.stride(0), a16.data_ptr(), a16.stride(0),
                                     b16.data_ptr(), b16.stride(0), b2_16.data_ptr(), b2_16.stride(0), r,
                                     dbias.data_ptr() if dbias is not None else None, _stream(a16.device))
    _lib.check(rc, "sam2b200_wgrad2")


def mlp_dh(dm16, w2_16, h16, scale=1.0, dbias=None):
    """dh = (dm @ W2) * (h > 0) * scale in one tcgen05 GEMM with the ReLU / hidden-dropout backward in its epilogue; dbias
    (fp32 [F], optional) += column sums of dh = the bias gradient of linear1, taken from the tile in registers."""
    r, f = h16.shape
    assert dm16.is_contiguous() and w2_16.is_contiguous() and h16.is_contiguous() and dm16.dtype == w2_16.dtype == h16.dtype == BF16
    assert dbias is None or (dbias.dtype == F32 and dbias.numel() == f and dbias.is_contiguous())
    dh = torch.empty_like(h16)
    rc = _lib.load().sam2b200_mlp_dh(dm16.data_ptr(), w2_16.data_ptr(), h16.data_ptr(), dh.data_ptr(),
                                     dbias.data_ptr() if dbias is not None else None, r, f, float(scale), _stream(h16.device))
    _lib.check(rc, "sam2b200_mlp_dh")
    return dh


def dropout_inplace_(x16, drop):
    p_, seed, site = _drop(drop)
    if p_ > 0.0:
        rc = _lib.load().sam2b200_dropout_inplace(x16.data_ptr(), x16.numel(), p_, seed, site, _stream(x16.device))
        _lib.check(rc, "sam2b200_dropout_inplace")


def colsum_bf16(x16, colsum):
    _colsum(2, None, x16, None, colsum, x16.shape[0], x16.shape[1])


def permute_rows(a, a2, alpha2, bsz, length, inverse, out_dtype=F32, second=False, scale=1.0):
    """Seq-first [L, B, C] <-> batch-first [B, L, C] rows with fused add / scale / cast (sam2b200_permute_rows).
    Returns out (and out2 = the same from `a` alone if `second`), shaped [B*L, C] (forward) or [L, B, C] (inverse)."""
    c = a.shape[-1]
    a = a.contiguous()
    if a.dtype != F32:
        a = a.float()
    if a2 is not None:
        a2 = a2.contiguous()
        if a2.dtype != F32:
            a2 = a2.float()
    shape = (length, bsz, c) if inverse else (bsz * length, c)
    out = torch.empty(shape, dtype=out_dtype, device=a.device)
    out2 = torch.empty(shape, dtype=out_dtype, device=a.device) if second else None
    rc = _lib.load().sam2b200_permute_rows(a.data_ptr(), a2.data_ptr() if a2 is not None else None, float(alpha2), out.data_ptr(),
                                           out2.data_ptr() if out2 is not None else None, int(out_dtype == BF16), bsz, length, c,
                                           int(inverse), float(scale), _stream(a.device))
    _lib.check(rc, "sam2b200_permute_rows")
    return (out, out2) if second else out


def _mm32(a, b):
    """fp32 result of a bf16 x bf16 product (weight gradients)."""
    return torch.mm(a, b, out_dtype=F32)


NO_DIRECT_GRADS = bool(os.environ.get("SAM2B200_NO_DIRECT_GRADS"))   # A/B switch: return gradients to autograd instead


def direct_grads_possible(bucket, params) -> bool:
    """True iff every parameter's .grad is (still) its fp32 view of `bucket`: the backward may then accumulate into
    the bucket itself and hide the parameters from autograd (they are passed detached)."""
    return (bucket is not None and not NO_DIRECT_GRADS and torch.is_grad_enabled()
            and all(p.requires_grad and bucket.owns(p) for p in params))


# The fused projection kernel (csrc/proj.cu: bias + RoPE in the GEMM epilogue) is OPT-IN: measured on B200 at cfg2 it ties
# cuBLAS addmm + the RoPE pass inside the graph-replayed step (56.9 vs 56.6 ms) and loses for the K = 64 memory-bank
# projections (profiles/r1_proj_rope_bench.txt).  SAM2B200_PROJ_KERNEL=1 routes the K = 256 projections through it,
# SAM2B200_PROJ_KERNEL_K64=1 the K = 64 ones as well.
# The folded projection's parameter gradients (four [256 x 64]-sized fp32 products + accumulations) in one launch of sam2b200_fold_grads
# instead of nine tiny ATen / cuBLAS launches: correct (the stack parity tests pass with it) but 0.2-0.3 ms per cfg2 step SLOWER in two A/B
# runs (its 256 blocks take SMs from the main stream; the tiny launches hide on the side stream) -> opt-in, SAM2B200_FOLD_KERNEL=1.
NO_FOLD_KERNEL = not bool(os.environ.get("SAM2B200_FOLD_KERNEL"))
NO_FOLD = bool(os.environ.get("SAM2B200_NO_FOLD"))  # A/B switch: v_proj and out_proj of the raw-memory cross-attention as two GEMMs
NO_V64 = bool(os.environ.get("SAM2B200_NO_V64"))    # A/B switch: cross-attention on the projected 256-d values (with the dV kernel)
NO_PROJ_KERNEL = not bool(os.environ.get("SAM2B200_PROJ_KERNEL"))
PROJ_KERNEL_K64 = bool(os.environ.get("SAM2B200_PROJ_KERNEL_K64"))
# q / k / v bias gradients: column sums of dq / dk / dv.  Default: the attention kernels' gradient epilogues accumulate them
# (31-shuffle butterfly + fp32 atomics per 32 x 32 block, ~6 % of the backward kernels' time).  SAM2B200_SIDE_STREAM_BIAS=1 = a
# separate column-sum pass on the side stream instead: the attention backward gets 1.0 ms faster at cfg2 (roofline 0.465 ->
# 0.497) but the STEP gets 0.5 ms slower -- both streams together saturate the GPU, so re-reading the gradients (116 MB for the
# cross-attention keys) costs more than the epilogue arithmetic (profiles/r2_bias_gradient_ab.txt).  Not the default.
EPILOGUE_BIAS = not bool(os.environ.get("SAM2B200_SIDE_STREAM_BIAS"))
# Round 2, later: wherever the weight gradient of the same projection goes through sam2b200_wgrad, the bias gradient comes out of
# THAT kernel (16 extra accumulator columns against a constant operand of ones: no extra pass, no extra launch) and the attention
# epilogue skips its column sums.  SAM2B200_NO_WGRAD_BIAS=1 restores the epilogue sums everywhere.
WGRAD_BIAS = not bool(os.environ.get("SAM2B200_NO_WGRAD_BIAS"))


def bias_grad_(gbias, dx16):
    """gbias [C] fp32 += column sums of dx16 [R, C] bf16 (C = 256 or 768; row stride may exceed C): the two-stage
    deterministic column-sum kernel of csrc/glue.cu (HBM bound, ~6 us for a [32 256 x 256] gradient)."""
    assert dx16.dtype == BF16 and dx16.stride(1) == 1
    _colsum(2, None, dx16, None, gbias, dx16.shape[0], dx16.shape[1], ld=dx16.stride(0))


# Weight gradients with 64 or <= 768 x 256 outputs go to sam2b200_wgrad (own split-K tcgen05 kernel, in-place fp32 accumulation: 10.5
# vs 14.9 us for [256 x 256], 17.6 vs 20.8 us for the stacked q|k|v, 29 vs 35 us for the memory-key projection); the two MLP weights
# ([256 x 2048], [2048 x 256]) stay on cuBLAS, which is 5-15 % faster there (profiles/r2_wgrad_bench.txt).  A/B: SAM2B200_NO_WGRAD=1.
NO_WGRAD = bool(os.environ.get("SAM2B200_NO_WGRAD"))
NO_WGRAD2 = bool(os.environ.get("SAM2B200_NO_WGRAD2"))   # A/B: separate launches for the key weight gradient and the bank-segment sums


WGRAD_MLP = bool(os.environ.get("SAM2B200_WGRAD_MLP"))   # A/B: the two MLP weight gradients on sam2b200_wgrad as well


def _wgrad_ok(mo, no):
    if WGRAD_MLP and not NO_WGRAD and ((mo, no) == (256, 2048) or (mo, no) == (2048, 256)):
        return True
    return (not NO_WGRAD) and mo % 256 == 0 and (no == 64 or (no == 256 and mo <= 768))


# Dense projections without a LayerNorm in front (linear2, the memory-key / value projections, the un-fused out_proj fall-backs) and
# every input-gradient GEMM dX = dY W run on sam2b200_gemm (csrc/gemm.cu: resident CTAs, SS-mode tcgen05, bias / RoPE epilogue).
# A/B: SAM2B200_NO_GEMM=1 sends them back to cuBLAS (torch.addmm / mm) and the stand-alone RoPE pass.
NO_GEMM = bool(os.environ.get("SAM2B200_NO_GEMM"))


def _gemm_ok(a, k, no):
    return (not NO_GEMM) and no in (256, 64) and k % 64 == 0 and a.dim() == 2 and a.stride(1) == 1 and a.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0


def linear_fwd(a16, w16, bias32, bias16):
    """a16 @ w16^T + bias (nn.Linear forward, w16 [No, K]); the fp32 master bias feeds our kernel, the bf16 mirror cuBLAS."""
    if _gemm_ok(a16, w16.shape[1], w16.shape[0]) and w16.is_contiguous():
        return gemm(a16, w16, bias=bias32)
    return torch.addmm(bias16, a16, w16.t())


def linear_dgrad_dot(dy16, w16, rows32):
    """(dx, dot): dx = dy16 @ w16 ([R, 64], bf16) and dot[r] = sum_c dx[r, c] * rows32[r, c] (fp32) -- the attention backward's
    Delta = rowsum(dO' o out64) comes out of the epilogue of the GEMM that produces dO' (no cast / multiply / row-sum passes)."""
    if _gemm_ok(dy16, w16.shape[0], w16.shape[1]) and w16.is_contiguous() and w16.shape[1] == 64 and rows32.is_contiguous():
        return gemm(dy16, w16, nn=True, dot_rows=rows32)
    dx = torch.mm(dy16, w16)
    return dx, (dx.float() * rows32.view(dx.shape)).sum(-1)


def linear_dgrad(dy16, w16):
    """dy16 @ w16 (nn.Linear input gradient, w16 [K = out_features, No = in_features])."""
    if _gemm_ok(dy16, w16.shape[0], w16.shape[1]) and w16.is_contiguous():
        return gemm(dy16, w16, nn=True)
    return torch.mm(dy16, w16)


NO_FUSED_OUT_PROJ = bool(os.environ.get("SAM2B200_NO_FUSED_OUT_PROJ"))   # A/B switch: out_proj as a separate cuBLAS addmm after the attention kernel
# Heads of the pre-norm blocks: default = ln_fwd + sam2b200_gemm_ex (two kernels); SAM2B200_LNPROJ_FUSED=1 = the one-kernel sam2b200_ln_proj.
LNPROJ_FUSED = bool(os.environ.get("SAM2B200_LNPROJ_FUSED"))
NO_LNPROJ = bool(os.environ.get("SAM2B200_NO_LNPROJ"))     # A/B switch: ln_fwd + cuBLAS addmm + RoPE pass instead of sam2b200_ln_proj
NO_MLP_KERNEL = bool(os.environ.get("SAM2B200_NO_MLP_KERNEL"))     # A/B switch: cuBLAS GEMM + separate ReLU-backward pass
NO_LN_FOLD_ON_SIDE = bool(os.environ.get("SAM2B200_NO_LN_FOLD_ON_SIDE"))   # A/B: LayerNorm-backward parameter folds on the main stream
NO_SIDE_STREAM = bool(os.environ.get("SAM2B200_NO_SIDE_STREAM"))   # A/B switch: everything on one stream
_SIDE_STREAMS = {}


class _SideStream:
    """Second CUDA stream for work that is off the critical path of the stack:
      * forward: the cross-attention K/V projections + RoPE of all layers depend only on the memory bank, so they run
        ahead of the layers (HBM-bound) while the main stream is in the tensor-core-bound self-attention;
      * backward: weight / bias gradients only feed the optimizer, so their long-K split-K GEMMs and column sums fill
        the SMs the attention kernels leave idle (wave tails) instead of sitting between them.
    Tensors produced on one stream and consumed on the other are kept alive until `join()`; fork / join use events,
    which CUDA-graph capture records as graph dependencies."""

    def __init__(self, dev):
        self.enabled = not NO_SIDE_STREAM
        self.main = torch.cuda.current_stream(dev)
        if self.enabled:
            key = (dev.index, self.main.cuda_stream)
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
            self.side = _SIDE_STREAMS[key]
        self.keep = []
        self.used = False

    def run(self, fn, *keep):
        """Enqueue fn() on the side stream after everything enqueued so far on the main stream."""
        if not self.enabled:
            return fn()
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            out = fn()
        self.keep.extend(keep)
        self.used = True
        return out

    def event(self):
        if not self.enabled:
            return None
        ev = torch.cuda.Event()
        ev.record(self.side)
        return ev

    def wait(self, ev):
        if ev is not None:
            self.main.wait_event(ev)

    def join(self):
        if self.enabled and self.used:
            self.main.wait_stream(self.side)
        self.keep.clear()
        self.used = False


class WeightMirror:
    """Persistent bf16 copies of the fp32 master parameters in ONE flat buffer (static addresses, so CUDA graphs can
    read them), refreshed by a single multi-tensor copy when a parameter was updated in place (tensor._version
    changes on optimizer.step / copy_ / load_state_dict): one cast per optimizer step instead of 106 small cast
    kernels in every forward and every backward (18 calls per clip at T = 10).  `refresh()` is called on the host
    before the kernels of a forward are enqueued -- never inside a graph capture."""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        # layout: per layer the q/k/v self-attention weights back to back (the backward multiplies by the stacked
        # [768, 256] matrix), everything else in parameter order; every numel is a multiple of 64 -> aligned views
        nl = (len(self.params) - 2) // _NPL if (len(self.params) - 2) % _NPL == 0 else 0
        first = []
        for l in range(nl):
            first += [l * _NPL + _LAYER_KEYS.index(k) for k in ("sa.q.w", "sa.k.w", "sa.v.w")]
        for l in range(nl):     # ... and their biases (the fused q|k|v projection adds a stacked [768] bias)
            first += [l * _NPL + _LAYER_KEYS.index(k) for k in ("sa.q.b", "sa.k.b", "sa.v.b")]
        order = first + [i for i in range(len(self.params)) if i not in set(first)]
        sizes = [-(-self.params[i].numel() // 64) * 64 for i in order]
        self.flat = torch.empty(sum(sizes), dtype=BF16, device=dev)
        self.views, off = [None] * len(self.params), 0
        self.qkv = []
        for i, n in zip(order, sizes):
            self.views[i] = self.flat[off:off + self.params[i].numel()].view_as(self.params[i])
            off += n
        self.qkv_bias = []
        for l in range(nl):
            v0 = self.views[first[3 * l]]
            o0 = v0.storage_offset()
            self.qkv.append(self.flat[o0:o0 + 3 * v0.numel()].view(3 * v0.shape[0], v0.shape[1]))
            b0 = self.views[first[3 * nl + 3 * l]]
            self.qkv_bias.append(self.flat[b0.storage_offset():b0.storage_offset() + 3 * b0.numel()])
        # raw-memory cross-attention: out_proj(v_proj(.)) folded per layer -- Wo Wv [256, 64], Wo bv + bo and Wo bv [256],
        # recomputed from the fp32 masters on refresh (static addresses, like the mirror itself)
        self.nl = nl
        self.w_eff = [torch.empty(256, 64, dtype=BF16, device=dev) for _ in range(nl)]
        self.b_eff = [torch.empty(256, dtype=BF16, device=dev) for _ in range(nl)]
        self.wobv = [torch.empty(256, dtype=BF16, device=dev) for _ in range(nl)]
        # fp32 copies for the attention kernels' fused output projection (bias added on the fp32 accumulator)
        self.b_eff32 = [torch.empty(256, dtype=F32, device=dev) for _ in range(nl)]
        self.qkv_bias32 = [torch.empty(768, dtype=F32, device=dev) for _ in range(nl)]    # stacked fp32 q|k|v bias (epilogue of gemm_ex)
        self.wk_all32 = torch.empty(max(nl, 1) * 256, 64, dtype=F32, device=dev)      # cross-attention key weights of all layers, stacked (bank position gradients)
        self.wobv32 = [torch.empty(256, dtype=F32, device=dev) for _ in range(nl)]
        self.versions = None
        self.device = dev

    def matches(self, params) -> bool:
        return (len(params) == len(self.params) and all(a is b for a, b in zip(params, self.params))
                and params[0].device == self.device)

    def invalidate(self):
        """Force the next refresh() to re-cast (call after updates the version counter cannot see; an
        ``optimizer.register_step_post_hook(lambda *_: mirror.invalidate())`` covers third-party optimizers)."""
        self.versions = None

    def refresh(self):
        # (version counter, storage address) per parameter: `p.data = new` / `p.data.copy_()` style updates keep the
        # version but usually move or at least can be announced through invalidate(); in debug mode one slice per
        # parameter is compared with its master as well
        v = [(p._version, p.data_ptr()) for p in self.params]
        if DEBUG_MIRROR and v == self.versions and not torch.cuda.is_current_stream_capturing():
            for view, p in zip(self.views, self.params):
                if not torch.equal(view.flatten()[:64], p.detach().flatten()[:64].to(BF16)):
                    raise RuntimeError("WeightMirror is stale: a parameter was updated through .data without bumping its "
                                       "version counter; call fused_stack.invalidate_mirrors(module) after such updates")
        if v != self.versions:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("WeightMirror.refresh() inside a CUDA graph capture: refresh before capturing")
            with torch.no_grad():
                torch._foreach_copy_(self.views, [p.detach() for p in self.params])
                for l in range(self.nl):
                    for j, kb in enumerate(("sa.q.b", "sa.k.b", "sa.v.b")):
                        self.qkv_bias32[l][256 * j:256 * (j + 1)].copy_(self.params[l * _NPL + _LAYER_KEYS.index(kb)].detach())
                    wo, bo, wv, bv = (self.params[l * _NPL + _LAYER_KEYS.index(k)].detach()
                                      for k in ("ca.o.w", "ca.o.b", "ca.v.w", "ca.v.b"))
                    wk = self.params[l * _NPL + _LAYER_KEYS.index("ca.k.w")].detach()
                    if wk.shape == (256, 64):
                        self.wk_all32[256 * l:256 * (l + 1)].copy_(wk)
                    if wv.shape != (256, 64):
                        continue
                    self.w_eff[l].copy_(torch.mm(wo, wv))
                    wobv = torch.mv(wo, bv)
                    self.wobv[l].copy_(wobv)
                    self.b_eff[l].copy_(wobv + bo)
                    self.wobv32[l].copy_(wobv)
                    self.b_eff32[l].copy_(wobv + bo)
            self.versions = v
        return self.views


DEBUG_MIRROR = bool(os.environ.get("SAM2B200_DEBUG_MIRROR"))
# the mirror lives ON the first parameter tensor (an attribute of the nn.Parameter object): it is dropped together with
# the model instead of accumulating in a global dict
_MIRROR_ATTR = "_sam2b200_weight_mirror"


def weight_mirror(params) -> WeightMirror:
    m = getattr(params[0], _MIRROR_ATTR, None)
    if m is None or not m.matches(params):
        m = WeightMirror(params)
        setattr(params[0], _MIRROR_ATTR, m)
    return m


def invalidate_mirrors(module) -> None:
    """Mark the bf16 weight mirror of `module` (a MemoryAttention) stale: the next forward re-casts every parameter and
    recomputes the folded projections.  Needed only after updates made through ``p.data`` (EMA swaps,
    ``vector_to_parameters``, optimizers that bypass the version counter)."""
    params = [p for _, p in module.named_parameters()]
    m = getattr(params[0], _MIRROR_ATTR, None) if params else None
    if m is not None:
        m.invalidate()


def bf16_params(params):
    """bf16 views of the parameters (see WeightMirror); keyed by the identity of the first parameter."""
    return weight_mirror(params).refresh()


_SEG_IND = {}


def segment_indicator(dev, b, m, n_slots, hw, n_ptrs):
    """[B * M, 64] bf16 one-hot rows: column j marks the bank segment of the row -- memory slot j (tokens [j hw, (j + 1) hw)) for
    j < n_slots, pointer j - n_slots behind them.  S = dk^T . Ind (one sam2b200_wgrad launch) gives the key gradient summed over
    objects and over the tokens of every segment, which is all the bank's position tensors need:
    d tpos[j] = sum_layers S_l[:, j]^T Wk_l  (sum over rows commutes with the projection).  Cached per bank shape."""
    key = (dev.index, b, m, n_slots, hw, n_ptrs)
    ind = _SEG_IND.get(key)
    if ind is None:
        if n_slots + n_ptrs > 64:
            raise _lib.Sam2B200Error("packed bank with more than 64 segments")
        sp = n_slots * hw
        t = torch.arange(m, device=dev)
        per = (m - sp) // n_ptrs if n_ptrs else 1
        seg = torch.where(t < sp, t // max(hw, 1), (n_slots + (t - sp) // max(per, 1)) if n_ptrs else torch.full_like(t, 63))
        ind = torch.nn.functional.one_hot(seg, 64).to(BF16).unsqueeze(0).expand(b, m, 64).reshape(b * m, 64).contiguous()
        while len(_SEG_IND) >= 64:          # bounded cache; a backward that uses an indicator keeps its own reference (ctx), so an
            _SEG_IND.pop(next(iter(_SEG_IND)))   # evicted entry that a captured CUDA graph still reads stays alive with that graph
        _SEG_IND[key] = ind
    return ind


class MemoryAttentionStackFn(torch.autograd.Function):
    """out[N,B,256] = MemoryAttention(curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens)."""

    @staticmethod
    def forward(ctx, meta, curr, curr_pos, memory, memory_pos, bank_tpos, bank_objpos, *params):
        # bank_tpos [n_slots, 64] / bank_objpos [n_ptrs, 64] (packed bank only, else None): the differentiable tensors the key
        # source memk depends on; their gradients are reductions of the fp32 d memk over tokens and objects (backward)
        # direct-gradient mode: params are detached aliases and the LAST tensor is a 1-element leaf with
        # requires_grad=True (the "grad anchor") whose only job is to make the output require grad, so that the
        # backward -- which accumulates the parameter gradients itself -- always runs
        ctx.has_anchor = bool(meta.get("direct"))
        if ctx.has_anchor:
            params = params[:-1]
        nl, p_excl, table, pos_at_input = meta["num_layers"], meta["num_k_exclude_rope"], meta["table"], meta["pos_enc_at_input"]
        n, b, d = curr.shape
        # packed bank (memory_bank.PackedBank): `memory` IS memk = bf16(memory + pos) and `memory_pos` IS memv = bf16(memory),
        # both [B, M, 64] batch-first -- written once by sam2b200_bank_gather_packed, consumed here without a re-pack
        packed = bool(meta.get("packed"))
        m = memory.shape[1] if packed else memory.shape[0]
        r, rm = b * n, b * m
        scale = 1.0 / math.sqrt(d)
        n_rope_k = m - p_excl
        # ---- pack inputs once per call (memory_attention.py:140-148; the reference re-adds pos per layer)
        # one pass each: x = curr + 0.1 curr_pos as batch-first fp32 rows; memk = bf16(memory + pos), memv = bf16(memory)
        x = permute_rows(curr, curr_pos if (pos_at_input and curr_pos is not None) else None, 0.1, b, n, False)
        if packed:
            assert memory.shape == (b, m, 64) and memory_pos.shape == (b, m, 64) and memory.dtype == BF16 and memory_pos.dtype == BF16
            memk, memv = memory.contiguous().view(rm, 64), memory_pos.contiguous().view(rm, 64)
        else:
            memk, memv = permute_rows(memory, memory_pos, 1.0, b, m, False, out_dtype=BF16, second=True)
        masters = meta.get("master_params") or list(params)   # the nn.Parameters (params may be detached aliases)
        wb = bf16_params(masters)
        saved: List[torch.Tensor] = []
        res = None
        side = _SideStream(x.device)
        kv_ready = []
        # train-mode dropout (memory_attention.py:58-99, transformer.py:304-306): dr = dict(seed, p_res, p_sa, p_ca) | None.
        # Site ids: layer * 8 + {0 self-attn probabilities, 1 cross-attn probabilities, 2 dropout1, 3 dropout2,
        # 4 MLP hidden, 5 dropout3}; the masks are regenerated from (seed, site) in the backward.
        dr = meta.get("dropout")

        def dsite(p_key, l, k):
            return (dr[p_key], dr["seed"], l * 8 + k) if dr is not None and dr[p_key] > 0.0 else None

        mirror = weight_mirror(masters)
        # Cross-attention on the raw 64-d memory features (sam2b200_attn_fwd_v64): softmax rows sum to 1, so
        # softmax(.) (mem Wv^T + bv) = (softmax(.) mem) Wv^T + bv -- v_proj moves from the [B M, 64] bank to the [B N, 64]
        # result, the [B, M, 256] value tensor and the dV kernel disappear.  Needs: no gradient w.r.t. `memory` (the bank
        # is detached in training, sam2model.py:345-358) and enough (query block, object) CTAs to fill the GPU without
        # split-KV.  With attention-probability dropout the rows of the dropped matrix do not sum to 1: the kernels then
        # also return those row sums (factor of the value bias) and take the per-query constant dO . bv in the backward
        # (folded projections only).
        ca_drop = dr is not None and dr["p_ca"] > 0.0
        # gradient w.r.t. the VALUE source: input 3 (memory) in the reference layout; a packed bank's features are constants
        need_v_grad = False if packed else ctx.needs_input_grad[3]
        v64 = (not NO_V64 and not (ca_drop and NO_FOLD) and not need_v_grad and b * ((n + 127) // 128) >= 64)

        def project_memory():        # cross-attention keys / values of every layer: functions of the bank only
            for l in range(nl):
                W = dict(zip(_LAYER_KEYS, wb[l * _NPL:(l + 1) * _NPL]))
                Pm = dict(zip(_LAYER_KEYS, params[l * _NPL:(l + 1) * _NPL]))
                if not NO_GEMM and not PROJ_KERNEL_K64 and _gemm_ok(memk, 64, d):
                    # K = 64 on resident CTAs (one k-slice per tile, the epilogue of tile i under the loads of tile i + 1), bias and
                    # rotation on the fp32 accumulator: no un-rotated memory keys in HBM, no RoPE pass
                    k2_rot = gemm(memk, W["ca.k.w"], bias=Pm["ca.k.b"], table=table, rows_per_item=m, n_rope_rows=n_rope_k).view(b, m, d)
                    v2 = None if v64 else gemm(memv, W["ca.v.w"], bias=Pm["ca.v.b"])
                elif not PROJ_KERNEL_K64:
                    # A/B (SAM2B200_NO_GEMM=1): cuBLAS + the RoPE pass; the one-tile-per-CTA proj_rope kernel below is slower at K = 64
                    # (1777 CTAs of 64 KB output each are dominated by per-CTA set-up: 129 vs 80 us, profiles/r1_proj_rope_bench.txt)
                    k2 = torch.addmm(W["ca.k.b"], memk, W["ca.k.w"].t())
                    v2 = None if v64 else torch.addmm(W["ca.v.b"], memv, W["ca.v.w"].t())
                    k2_rot = rope_apply(k2.view(b, m, d), table, n_rope_k)
                else:
                    k2_rot = proj_rope(memk, W["ca.k.w"], W["ca.k.b"], 1, table, 1, m, n_rope_k)[0].view(b, m, d)
                    v2 = None if v64 else proj_rope(memv, W["ca.v.w"], W["ca.v.b"], 1)[0]
                kv_ready.append((k2_rot, v2, side.event()))

        side.run(project_memory, memk, memv)
        for l in range(nl):
            P = dict(zip(_LAYER_KEYS, params[l * _NPL:(l + 1) * _NPL]))
            W = dict(zip(_LAYER_KEYS, wb[l * _NPL:(l + 1) * _NPL]))
            # ---- self attention (memory_attention.py:58-64)
            if not NO_LNPROJ:
                # LayerNorm + q|k|v projection + bias + RoPE(q, k) in one kernel: no un-rotated q / k in HBM
                (q_rot, k_rot, v), y1, x, mean1, rstd1 = block_head(
                    x, res, P["n1.w"], P["n1.b"], mirror.qkv[l], mirror.qkv_bias[l], mirror.qkv_bias32[l], 3, table=table, rope_outs=2,
                    rows_per_item=n, n_rope_rows=n, drop_res=dsite("p_res", l - 1, 5) if l > 0 else None)
                q_rot, k_rot = q_rot.view(b, n, d), k_rot.view(b, n, d)
            else:
                y1, x, mean1, rstd1 = ln_fwd(x, res, P["n1.w"], P["n1.b"], drop=dsite("p_res", l - 1, 5) if l > 0 else None)
            if not NO_LNPROJ:
                pass
            elif NO_PROJ_KERNEL:
                q = torch.addmm(W["sa.q.b"], y1, W["sa.q.w"].t())
                k = torch.addmm(W["sa.k.b"], y1, W["sa.k.w"].t())
                v = torch.addmm(W["sa.v.b"], y1, W["sa.v.w"].t())
                q_rot = rope_apply(q.view(b, n, d), table, n)
                k_rot = rope_apply(k.view(b, n, d), table, n)
            else:   # one GEMM for q | k | v with bias, q and k rotated in its epilogue
                q_rot, k_rot, v = proj_rope(y1, mirror.qkv[l], mirror.qkv_bias[l], 3, table, 2, n, n)
                q_rot, k_rot = q_rot.view(b, n, d), k_rot.view(b, n, d)
            fuse_out = (not NO_FUSED_OUT_PROJ) and meta["nsplit"] in (0, 1) and b * ((n + 127) // 128) >= 100
            if fuse_out:   # out_proj inside the attention kernel's epilogue (transformer.py:308-309): no separate GEMM, o is not re-read
                o, o32, lse, sa = attn_fwd_proj(q_rot, k_rot, v.view(b, n, d), W["sa.o.w"], P["sa.o.b"], scale, drop=dsite("p_sa", l, 0))
                sa = sa.view(r, d)
            else:
                o, o32, lse = attn_fwd(q_rot, k_rot, v.view(b, n, d), scale, meta["nsplit"], drop=dsite("p_sa", l, 0))
                sa = linear_fwd(o.view(r, d), W["sa.o.w"], P["sa.o.b"], W["sa.o.b"])
            # ---- cross attention to the memory bank (memory_attention.py:66-81)
            if not NO_LNPROJ:
                (q2_rot,), y2, x1, mean2, rstd2 = block_head(x, sa, P["n2.w"], P["n2.b"], W["ca.q.w"], W["ca.q.b"], P["ca.q.b"], 1, table=table,
                                                             rope_outs=1, rows_per_item=n, n_rope_rows=n, drop_res=dsite("p_res", l, 2))
                q2_rot = q2_rot.view(b, n, d)
            else:
                y2, x1, mean2, rstd2 = ln_fwd(x, sa, P["n2.w"], P["n2.b"], drop=dsite("p_res", l, 2))
            if not NO_LNPROJ:
                pass
            elif NO_PROJ_KERNEL:
                q2 = torch.addmm(W["ca.q.b"], y2, W["ca.q.w"].t())
                q2_rot = rope_apply(q2.view(b, n, d), table, n)
            else:
                q2_rot = proj_rope(y2, W["ca.q.w"], W["ca.q.b"], 1, table, 1, n, n)[0].view(b, n, d)
            k2_rot, v2, ev = kv_ready[l]
            side.wait(ev)
            if v64 and not NO_FOLD and not NO_FUSED_OUT_PROJ:
                # raw-memory cross-attention with the folded out_proj(v_proj(.)) in the kernel's epilogue; the o2 slot of `saved`
                # holds the folded weight, the v2 / o2_32 slots out64 and its fp32 copy
                o2 = mirror.w_eff[l]
                ca_d = dsite("p_ca", l, 1)
                v2, o2_32, lse2, rs, ca = attn_fwd_v64_proj(q2_rot, k2_rot, memv.view(b, m, 64), o2,
                                                            P["ca.o.b"] if ca_d is not None else mirror.b_eff32[l],
                                                            mirror.wobv32[l] if ca_d is not None else None, scale, drop=ca_d)
                ca = ca.view(r, d)
            elif v64:
                # v2 / o2_32 slots of `saved` then hold out64 (bf16) and its fp32 copy
                v2, o2_32, lse2, rs = attn_fwd_v64(q2_rot, k2_rot, memv.view(b, m, 64), scale, drop=dsite("p_ca", l, 1))
                if NO_FOLD:
                    o2 = linear_fwd(v2.view(r, 64), W["ca.v.w"], P["ca.v.b"], W["ca.v.b"]).view(b, n, d)      # v_proj on the result
                    ca = linear_fwd(o2.view(r, d), W["ca.o.w"], P["ca.o.b"], W["ca.o.b"])
                else:
                    # out_proj(v_proj(out64)) = out64 (Wo Wv)^T + (Wo bv + bo): one [B N, 64] -> 256 GEMM; the o2 slot of
                    # `saved` holds the folded weight (fp32 product of the master weights, rounded once)
                    o2 = mirror.w_eff[l]                       # Wo Wv, refreshed with the weight mirror (once per optimizer step)
                    if rs is None:
                        ca = linear_fwd(v2.view(r, 64), o2, mirror.b_eff32[l], mirror.b_eff[l])
                    else:   # dropout: the value bias enters with the row sums of the dropped probabilities (rank-1 term)
                        ca = linear_fwd(v2.view(r, 64), o2, P["ca.o.b"], W["ca.o.b"])
                        ca.addr_(rs.view(r).to(BF16), mirror.wobv[l])
            else:
                rs = None
                o2, o2_32, lse2 = attn_fwd(q2_rot, k2_rot, v2.view(b, m, d), scale, meta["nsplit"], drop=dsite("p_ca", l, 1))
                ca = linear_fwd(o2.view(r, d), W["ca.o.w"], P["ca.o.b"], W["ca.o.b"])
            # ---- MLP (memory_attention.py:95-98)
            if not NO_LNPROJ:   # LayerNorm + linear1 + bias + ReLU (+ hidden dropout) in one kernel
                (h,), y3, x2, mean3, rstd3 = block_head(x1, ca, P["n3.w"], P["n3.b"], W["l1.w"], W["l1.b"], P["l1.b"], 1, out_width=2048, relu=True,
                                                        drop_res=dsite("p_res", l, 3), drop_out=dsite("p_res", l, 4))
            else:
                y3, x2, mean3, rstd3 = ln_fwd(x1, ca, P["n3.w"], P["n3.b"], drop=dsite("p_res", l, 3))
                h = torch._addmm_activation(W["l1.b"], y3, W["l1.w"].t(), use_gelu=False)  # bias + ReLU epilogue
                dropout_inplace_(h, dsite("p_res", l, 4))
            mlp = linear_fwd(h, W["l2.w"], P["l2.b"], W["l2.b"])
            saved += [x, mean1, rstd1, y1, q_rot, k_rot, v, o, o32, lse,
                      x1, mean2, rstd2, y2, q2_rot, k2_rot, v2, o2, o2_32, lse2,
                      x2, mean3, rstd3, y3, h, rs if rs is not None else lse2[:0]]
            x, res = x2, mlp
        gamma_f, beta_f = params[nl * _NPL], params[nl * _NPL + 1]
        out, x_fin, mean_f, rstd_f = ln_fwd(x, res, gamma_f, beta_f, want_f32_seq_first=(b, n), drop=dsite("p_res", nl - 1, 5))
        side.join()
        saved += [x_fin, mean_f, rstd_f, memk, memv, table]
        ctx.save_for_backward(*saved, *params)
        ctx.n_saved = len(saved)
        ctx.meta = dict(nl=nl, n=n, b=b, m=m, scale=scale, n_rope_k=n_rope_k, pos_at_input=pos_at_input,
                        has_pos=curr_pos is not None, bucket=meta.get("bucket"), masters=masters,
                        direct=bool(meta.get("direct")), dropout=dr, v64=v64, fold=v64 and not NO_FOLD, packed=packed,
                        bank_slots=int(meta.get("bank_slots", 0)), bank_hw=int(meta.get("bank_hw", 0)),
                        bank_ptrs=int(bank_objpos.shape[0]) if bank_objpos is not None else 0)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        mt = ctx.meta
        nl, n, b, m, scale, n_rope_k = mt["nl"], mt["n"], mt["b"], mt["m"], mt["scale"], mt["n_rope_k"]
        d = 256
        r, rm = b * n, b * m
        saved = ctx.saved_tensors[:ctx.n_saved]
        params = ctx.saved_tensors[ctx.n_saved:]
        x_fin, mean_f, rstd_f, memk, memv, table = saved[-6:]
        dev = x_fin.device
        need_curr, need_pos, need_mem, need_mpos = ctx.needs_input_grad[1:5]
        packed = mt.get("packed", False)
        need_tpos, need_objpos = ctx.needs_input_grad[5:7]
        if packed:      # memk / memv are constants; the key-source gradient is wanted iff the bank's position tensors ask for it
            need_mem, need_mpos = False, bool(need_tpos or need_objpos)      # below: need_mpos = "dmemk wanted"
        need_memgrad = need_mem or need_mpos
        masters = mt["masters"]
        wb = bf16_params(masters)
        wqkv_all = weight_mirror(masters).qkv
        # Parameter gradients.  With a GradBucket attached (ddp.attach_grad_bucket) whose views are still the
        # parameters' .grad, every kernel / GEMM ACCUMULATES straight into the bucket (beta = 1) and autograd gets
        # None for the parameters: no fp32 temporaries, no 106 AccumulateGrad add_ kernels per backward.  Otherwise
        # vector gradients accumulate into one zeroed scratch and matrix gradients are returned from the GEMMs.
        bucket = mt.get("bucket")
        direct = mt["direct"]
        grads = [None] * len(params)
        if direct:
            if not all(bucket.owns(p) for p in masters):
                raise RuntimeError("a parameter's .grad was detached from the GradBucket between forward and backward "
                                   "(e.g. zero_grad(set_to_none=True)); use bucket.zero() or re-attach the bucket")
            gv = [p.grad for p in masters]
        else:
            gv = [None] * len(params)
            vec = [i for i, p in enumerate(params) if p.dim() == 1]
            flat = torch.zeros(sum(params[i].numel() for i in vec), dtype=F32, device=dev)   # one memset for all vectors
            off = 0
            for i in vec:
                gv[i] = grads[i] = flat[off:off + params[i].numel()]
                off += params[i].numel()

        side = _SideStream(dev)
        dr = mt["dropout"]

        def dsite(p_key, l, k):
            return (dr[p_key], dr["seed"], l * 8 + k) if dr is not None and dr[p_key] > 0.0 else None
        relu_scale = 1.0 / (1.0 - dr["p_res"]) if dr is not None and dr["p_res"] > 0.0 else 1.0

        def acc_w(i, a_t, bmat, bias=None):     # grad[i] (+)= a_t @ bmat (bf16 x bf16 -> fp32); bias grad (+)= colsum(a_t^T)
            def work():
                own = _wgrad_ok(a_t.shape[0], bmat.shape[1]) and a_t.stride(0) == 1
                fused_b = bias if (own and WGRAD_BIAS) else None
                if bias is not None and fused_b is None:
                    colsum_bf16(a_t.t(), bias)
                if own:
                    if direct:
                        wgrad_(gv[i], a_t.t(), bmat, fused_b)
                    else:
                        grads[i] = wgrad_(torch.zeros((a_t.shape[0], bmat.shape[1]), dtype=F32, device=bmat.device), a_t.t(), bmat, fused_b)
                elif direct:
                    torch.addmm(gv[i], a_t, bmat, out_dtype=F32, out=gv[i])
                else:
                    grads[i] = _mm32(a_t, bmat)
            side.run(work, a_t, bmat)
        grad_out = grad_out.contiguous().float()
        # every LayerNorm backward also emits the bf16 copy of its output gradient (operand of the next GEMMs) and the
        # bias gradient of the projection in front of it: no separate cast + column-sum pass
        last = (nl - 1) * _NPL
        g, g16 = ln_bwd(grad_out, x_fin, mean_f, rstd_f, params[nl * _NPL], None, gv[nl * _NPL], gv[nl * _NPL + 1],
                        seq_first=(b, n), dbias=gv[last + _LAYER_KEYS.index("l2.b")], drop=dsite("p_res", nl - 1, 5), side=side)
        # Packed bank: only the position tensors want a gradient, i.e. the key-source gradient summed over objects and over the
        # tokens of each bank segment -- S_l = dk_l^T . Ind ([256, 64] per layer, one wgrad launch on the side stream) instead of
        # the dense [B M, 64] fp32 gradient (a [B M, 256] x [256, 64] GEMM per layer, a 58 MB zero fill and three reductions).
        seg_mode = packed and need_memgrad and memk.shape[1] == 64 and _wgrad_ok(d, 64)
        seg_ind = segment_indicator(dev, b, m, int(mt["bank_slots"]), int(mt["bank_hw"]), int(mt["bank_ptrs"])) if seg_mode else None
        ctx._seg_ind_keepalive = seg_ind      # a CUDA-graph capture of this backward reads it at every replay
        seg_sum = torch.zeros((nl * d, 64), dtype=F32, device=dev) if seg_mode else None
        dmemk = torch.zeros((rm, memk.shape[1]), dtype=F32, device=dev) if (need_memgrad and not seg_mode) else None
        dmemv = torch.zeros((rm, memv.shape[1]), dtype=F32, device=dev) if need_mem else None
        per = 26
        for l in reversed(range(nl)):
            (x0, mean1, rstd1, y1, q_rot, k_rot, v, o, o32, lse, x1, mean2, rstd2, y2, q2_rot, k2_rot, v2, o2, o2_32,
             lse2, x2, mean3, rstd3, y3, h, rs) = saved[l * per:(l + 1) * per]
            base = l * _NPL
            ix = {k: base + i for i, k in enumerate(_LAYER_KEYS)}
            W = {k: wb[ix[k]] for k in _LAYER_KEYS}
            P = {k: params[ix[k]] for k in _LAYER_KEYS}
            # ---- MLP backward
            dm = g16
            acc_w(ix["l2.w"], dm.t(), h)
            if NO_MLP_KERNEL:
                dh = torch.mm(dm, W["l2.w"])
                relu_bwd_colsum_(dh, h, gv[ix["l1.b"]], scale=relu_scale)
                acc_w(ix["l1.w"], dh.t(), y3)
            else:   # ReLU / hidden-dropout backward fused into the GEMM; the bias gradient joins the weight gradient (side stream)
                dh = mlp_dh(dm, W["l2.w"], h, relu_scale, dbias=gv[ix["l1.b"]])     # + bias gradient of linear1 (column sums of dh)
                acc_w(ix["l1.w"], dh.t(), y3)
            dy3 = linear_dgrad(dh, W["l1.w"])
            fold = mt["fold"]
            g_bo = torch.zeros(d, dtype=F32, device=dev) if fold else gv[ix["ca.o.b"]]    # colsum(dca), needed on its own when folded
            g, dca = ln_bwd(dy3, x2, mean3, rstd3, P["n3.w"], g, gv[ix["n3.w"]], gv[ix["n3.b"]], dbias=g_bo,
                            drop=dsite("p_res", l, 3), side=side)
            # ---- cross attention backward
            if not fold:
                acc_w(ix["ca.o.w"], dca.t(), o2.view(r, d))
                do2 = linear_dgrad(dca, W["ca.o.w"])
            # conjugate RoPE and the q / k / v bias gradients (column sums) are fused into the gradient epilogues.
            # Only dQ is on the path of the residual-stream gradient: the key-side kernels (dV, dK) and everything
            # they feed (weight gradients, memory-bank gradients) go to the side stream.
            wb_q = WGRAD_BIAS and _wgrad_ok(d, d)          # q bias gradient from the weight-gradient kernel instead of the dQ epilogue
            if mt["v64"]:
                # saved v2 / o2_32 are out64 / its fp32 copy.  v_proj acted on out64: its gradients are two small GEMMs,
                # dout64 = dO Wv feeds the attention backward, Delta = rowsum(dout64 o out64); there is no dV.
                o64, o64_32 = v2, o2_32
                if fold:
                    # ca = out64 (Wo Wv)^T + (Wo bv + bo):  G = d/d(Wo Wv) = dca^T out64, g = d/d(Wo bv + bo) = colsum(dca);
                    # dWo = G Wv^T + g bv^T, dWv = Wo^T G, dbv = Wo^T g, dbo = g -- four [256 x 64]-sized fp32 products
                    do64, delta = linear_dgrad_dot(dca, o2, o64_32)             # o2 slot = folded weight [256, 64]
                    do64, delta = do64.view(b, n, 64), delta.view(b, n)
                    ca_drop = rs.numel() > 0
                    # with dropout the value bias entered as rowsum (x) Wo bv: its gradients use g_rs = dca^T rowsum
                    # instead of g = colsum(dca), and the per-query constant c = dca . (Wo bv) goes into dP and Delta
                    g_rs = _mm32(dca.t(), rs.view(r, 1).to(BF16)).view(d) if ca_drop else g_bo
                    dp_bias = _mm32(dca, weight_mirror(masters).wobv[l].view(d, 1)).view(b, n) if ca_drop else None

                    def fold_grads(dca=dca, o64=o64, g_bo=g_bo, g_rs=g_rs, ix=ix, P=P):
                        if _wgrad_ok(d, 64):
                            G = wgrad_(torch.zeros((d, 64), dtype=F32, device=dev), dca, o64.view(r, 64))
                        else:
                            G = _mm32(dca.t(), o64.view(r, 64))
                        if not NO_FOLD_KERNEL and all(P[k].dtype == F32 and P[k].is_contiguous() for k in ("ca.o.w", "ca.v.w", "ca.v.b")):
                            # the four small products + accumulations in ONE launch (sam2b200_fold_grads, csrc/glue.cu)
                            if not direct:
                                grads[ix["ca.o.w"]] = torch.zeros((d, d), dtype=F32, device=dev)
                                grads[ix["ca.v.w"]] = torch.zeros((d, 64), dtype=F32, device=dev)
                            d_wo = gv[ix["ca.o.w"]] if direct else grads[ix["ca.o.w"]]
                            d_wv = gv[ix["ca.v.w"]] if direct else grads[ix["ca.v.w"]]
                            rc = _lib.load().sam2b200_fold_grads(G.data_ptr(), g_bo.data_ptr(), g_rs.data_ptr(), P["ca.o.w"].data_ptr(),
                                                                 P["ca.v.w"].data_ptr(), P["ca.v.b"].data_ptr(), d_wo.data_ptr(), d_wv.data_ptr(),
                                                                 gv[ix["ca.o.b"]].data_ptr(), gv[ix["ca.v.b"]].data_ptr(), _stream(dev))
                            _lib.check(rc, "sam2b200_fold_grads")
                            return
                        d_wo = torch.addmm(torch.outer(g_rs, P["ca.v.b"]), G, P["ca.v.w"].t())
                        d_wv = torch.mm(P["ca.o.w"].t(), G)
                        gv[ix["ca.o.b"]].add_(g_bo)
                        gv[ix["ca.v.b"]].addmv_(P["ca.o.w"].t(), g_rs)
                        if direct:
                            gv[ix["ca.o.w"]].add_(d_wo)
                            gv[ix["ca.v.w"]].add_(d_wv)
                        else:
                            grads[ix["ca.o.w"]], grads[ix["ca.v.w"]] = d_wo, d_wv
                    side.run(fold_grads, dca, o64, g_bo, g_rs)
                else:
                    dp_bias = None
                    acc_w(ix["ca.v.w"], do2.t(), o64.view(r, 64), gv[ix["ca.v.b"]])
                    do64, delta = linear_dgrad_dot(do2, W["ca.v.w"], o64_32)
                    do64, delta = do64.view(b, n, 64), delta.view(b, n)
                if dp_bias is not None:
                    delta = torch.addcmul(delta, dp_bias, rs)
                kw = dict(table=table, n_rope_k=n_rope_k, grad_dtype=BF16, dp_bias=dp_bias, drop=dsite("p_ca", l, 1))
                args = (q2_rot, k2_rot, memv.view(b, m, 64), do64, lse2, delta, scale)

                def key_side(args=args, kw=kw, l=l, ix=ix, W=W):
                    wb = WGRAD_BIAS and _wgrad_ok(d, 64)       # bias gradient from the weight-gradient kernel
                    _, dk2 = attn_bwd_v64(*args, parts=4, dbias=(None, gv[ix["ca.k.b"]] if (EPILOGUE_BIAS and not wb) else None), **kw)
                    dk2 = dk2.view(rm, d)
                    if not EPILOGUE_BIAS and not wb:
                        bias_grad_(gv[ix["ca.k.b"]], dk2)
                    if _wgrad_ok(d, 64):
                        gb_k = gv[ix["ca.k.b"]] if wb else None
                        if not direct:
                            grads[ix["ca.k.w"]] = torch.zeros((d, 64), dtype=F32, device=dev)
                        gkw = gv[ix["ca.k.w"]] if direct else grads[ix["ca.k.w"]]
                        if seg_mode and not NO_WGRAD2:      # weight gradient + per-segment sums of dk in ONE pass over dk (116 MB at cfg2)
                            wgrad2_(gkw, seg_sum[l * d:(l + 1) * d], dk2, memk, seg_ind, gb_k)
                            return
                        wgrad_(gkw, dk2, memk, gb_k)
                    elif direct:
                        torch.addmm(gv[ix["ca.k.w"]], dk2.t(), memk, out_dtype=F32, out=gv[ix["ca.k.w"]])
                    else:
                        grads[ix["ca.k.w"]] = _mm32(dk2.t(), memk)
                    if seg_mode:
                        wgrad_(seg_sum[l * d:(l + 1) * d], dk2, seg_ind)
                    elif need_memgrad:
                        torch.addmm(dmemk, dk2, W["ca.k.w"], out_dtype=F32, out=dmemk)
                side.run(key_side, q2_rot, k2_rot, memv, do64, lse2, delta, dp_bias)
                dq2, _ = attn_bwd_v64(*args, parts=8, dbias=(gv[ix["ca.q.b"]] if (EPILOGUE_BIAS and not wb_q) else None, None), **kw)
            else:
                delta = torch.empty((b, n), dtype=F32, device=dev)
                args = (q2_rot, k2_rot, v2.view(b, m, d), None, o2_32, do2.view(b, n, d), lse2, scale)
                kw = dict(table=table, n_rope_k=n_rope_k, grad_dtype=BF16, delta=delta, drop=dsite("p_ca", l, 1))
                attn_bwd(*args, parts=1, **kw)                                              # Delta = rowsum(dO o O)

                def key_side(args=args, kw=kw, l=l, ix=ix, W=W):
                    eb = EPILOGUE_BIAS
                    _, dk2, dv2 = attn_bwd(*args, parts=2 | 4, dbias=(None, gv[ix["ca.k.b"]] if eb else None, gv[ix["ca.v.b"]] if eb else None), **kw)
                    dk2, dv2 = dk2.view(rm, d), dv2.view(rm, d)
                    if not eb:
                        bias_grad_(gv[ix["ca.k.b"]], dk2)
                        bias_grad_(gv[ix["ca.v.b"]], dv2)
                    if direct:
                        torch.addmm(gv[ix["ca.k.w"]], dk2.t(), memk, out_dtype=F32, out=gv[ix["ca.k.w"]])
                        torch.addmm(gv[ix["ca.v.w"]], dv2.t(), memv, out_dtype=F32, out=gv[ix["ca.v.w"]])
                    else:
                        grads[ix["ca.k.w"]] = _mm32(dk2.t(), memk)
                        grads[ix["ca.v.w"]] = _mm32(dv2.t(), memv)
                    if seg_mode:
                        wgrad_(seg_sum[l * d:(l + 1) * d], dk2, seg_ind)
                    elif need_memgrad:
                        torch.addmm(dmemk, dk2, W["ca.k.w"], out_dtype=F32, out=dmemk)
                    if need_mem:
                        torch.addmm(dmemv, dv2, W["ca.v.w"], out_dtype=F32, out=dmemv)
                side.run(key_side, *args[:3], o2_32, do2, lse2, delta)
                dq2, _, _ = attn_bwd(*args, parts=8, dbias=(gv[ix["ca.q.b"]] if (EPILOGUE_BIAS and not wb_q) else None, None, None), **kw)
            dq2 = dq2.view(r, d)
            dy2 = linear_dgrad(dq2, W["ca.q.w"])
            acc_w(ix["ca.q.w"], dq2.t(), y2, gv[ix["ca.q.b"]] if wb_q else None)
            if not EPILOGUE_BIAS and not wb_q:
                side.run(lambda dq2=dq2, gb=gv[ix["ca.q.b"]]: bias_grad_(gb, dq2), dq2)
            g, dsa = ln_bwd(dy2, x1, mean2, rstd2, P["n2.w"], g, gv[ix["n2.w"]], gv[ix["n2.b"]], dbias=gv[ix["sa.o.b"]],
                            drop=dsite("p_res", l, 2), side=side)
            # ---- self attention backward
            acc_w(ix["sa.o.w"], dsa.t(), o.view(r, d))
            do = linear_dgrad(dsa, W["sa.o.w"])
            dqkv = torch.empty((b, n, 3 * d), dtype=BF16, device=dev)   # [dq | dk | dv], written in place by the kernels
            qkv_w = [masters[ix[k]] for k in ("sa.q.w", "sa.k.w", "sa.v.w")]
            gw = bucket.span(qkv_w, (3 * d, d)) if direct else None
            gb3 = bucket.span([masters[ix[k]] for k in ("sa.q.b", "sa.k.b", "sa.v.b")], (3 * d,)) if direct else None
            wb_sa = WGRAD_BIAS and gw is not None and gb3 is not None and _wgrad_ok(3 * d, d)   # the three bias gradients from the stacked weight-gradient GEMM
            attn_bwd(q_rot, k_rot, v.view(b, n, d), None, o32, do.view(b, n, d), lse, scale, table=table, n_rope_k=n,
                     grad_dtype=BF16, dq=dqkv[:, :, :d], dk=dqkv[:, :, d:2 * d], dv=dqkv[:, :, 2 * d:],
                     dbias=(gv[ix["sa.q.b"]], gv[ix["sa.k.b"]], gv[ix["sa.v.b"]]) if (EPILOGUE_BIAS and not wb_sa) else (None, None, None),
                     drop=dsite("p_sa", l, 0))
            dqkv = dqkv.view(r, 3 * d)
            if not EPILOGUE_BIAS and not wb_sa:
                gb = gb3
                if gb is not None:      # the bucket lays the three biases out back to back: one [1 x R] . [R x 768] product
                    side.run(lambda dqkv=dqkv, gb=gb: bias_grad_(gb, dqkv), dqkv)
                else:
                    for j, kb in enumerate(("sa.q.b", "sa.k.b", "sa.v.b")):
                        side.run(lambda dx=dqkv[:, j * d:(j + 1) * d], g_=gv[ix[kb]]: bias_grad_(g_, dx), dqkv)
            if gw is not None:
                # the bucket lays the three projection weights out back to back: one [768, 256] weight-gradient GEMM
                if _wgrad_ok(3 * d, d):
                    side.run(lambda dqkv=dqkv, y1=y1, gw=gw, gb=(gb3 if wb_sa else None): wgrad_(gw, dqkv, y1, gb), dqkv, y1)
                else:
                    side.run(lambda dqkv=dqkv, y1=y1, gw=gw: torch.addmm(gw, dqkv.t(), y1, out_dtype=F32, out=gw), dqkv, y1)
            elif direct:
                for j, kw in enumerate(("sa.q.w", "sa.k.w", "sa.v.w")):
                    acc_w(ix[kw], dqkv[:, j * d:(j + 1) * d].t(), y1)
            else:
                dw = _mm32(dqkv.t(), y1)                     # [768, 256] = d(Wq | Wk | Wv) in one GEMM
                grads[ix["sa.q.w"]], grads[ix["sa.k.w"]], grads[ix["sa.v.w"]] = dw[:d], dw[d:2 * d], dw[2 * d:]
            dy1 = linear_dgrad(dqkv, wqkv_all[l])            # stacked [768, 256] weights: contraction over 768, fp32 accumulation
            if l > 0:
                g, g16 = ln_bwd(dy1, x0, mean1, rstd1, P["n1.w"], g, gv[ix["n1.w"]], gv[ix["n1.b"]],
                                dbias=gv[base - _NPL + _LAYER_KEYS.index("l2.b")], drop=dsite("p_res", l - 1, 5), side=side)
            else:
                g = ln_bwd(dy1, x0, mean1, rstd1, P["n1.w"], g, gv[ix["n1.w"]], gv[ix["n1.b"]], side=side)
        side.join()
        # ---- unpack input gradients
        d_curr = d_pos = d_mem = d_mpos = None
        if need_curr:
            d_curr = permute_rows(g, None, 0.0, b, n, True)
        if need_pos and mt["has_pos"] and mt["pos_at_input"]:
            d_pos = permute_rows(g, None, 0.0, b, n, True, scale=0.1)
        d_tpos = d_objpos = None
        if packed:
            # d maskmem_tpos_enc rows / d pointer positions = the fp32 key-source gradient summed over tokens and objects
            ns, hw = mt["bank_slots"], mt["bank_hw"]
            sp = ns * hw
            n_ptr = mt["bank_ptrs"]
            if seg_mode:      # [64 segments, nl * 256] x [nl * 256, 64]: sum over layers of S_l^T Wk_l (fp32 masters)
                d_seg = torch.mm(seg_sum.t(), weight_mirror(masters).wk_all32[:nl * d])
                if need_tpos and ns:
                    d_tpos = d_seg[:ns]
                if need_objpos and m > sp:
                    d_objpos = d_seg[ns:ns + n_ptr]
            else:
                dk3 = dmemk.view(b, m, 64) if need_mpos else None
                if need_tpos and ns:
                    d_tpos = dk3[:, :sp].reshape(b, ns, hw, 64).sum((0, 2))
                if need_objpos and m > sp:
                    d_objpos = dk3[:, sp:].reshape(b, n_ptr, (m - sp) // n_ptr, 64).sum((0, 2))
        else:
            if need_mpos:
                d_mpos = permute_rows(dmemk, None, 0.0, b, m, True)
            if need_mem:
                d_mem = permute_rows(dmemk, dmemv, 1.0, b, m, True)
        if not ctx.has_anchor:
            return (None, d_curr, d_pos, d_mem, d_mpos, d_tpos, d_objpos, *grads)
        # The anchor gets a (zero) gradient only when it is the ONLY input that requires grad: a backward whose
        # outputs are all None makes the autograd engine synchronise the capturing stream with the (uncaptured)
        # stream of the anchor's AccumulateGrad node and invalidates a CUDA-graph capture (scripts/probe_capture.py).
        lonely = not (need_curr or need_pos or need_mem or need_mpos)
        return (None, d_curr, d_pos, d_mem, d_mpos, d_tpos, d_objpos, *grads, torch.zeros(1, dtype=F32, device=dev) if lonely else None)
