"""``MemoryEncoder`` and its parts with the reference's constructor arguments, attribute names and state_dict keys
(sam2_video/model/modeling/memory_encoder.py:17-181; ``LayerNorm2d`` / ``DropPath``: sam2_utils.py:92-153;
``PositionEmbeddingSine``: position_encoding.py:16-130) -- SURVEY.md section 8f rank 3, the step that runs right after the
hot path on every frame (``SAM2Base._encode_new_memory``, sam2_base.py:715-769) and produces the features the memory bank
holds.

Internally everything is channels-last / token-major (the layout of the bank and of the attention path).  What runs where:
  * LayerNorm2d + GELU after every strided convolution of ``MaskDownSampler``: ONE own kernel per stage (csrc/memenc.cu),
    forward and backward;
  * the 7 x 7 depth-wise convolution of ``CXBlock``: own kernels (forward, data gradient, weight gradient);
  * LayerNorm + ``pwconv1`` of ``CXBlock``: ``sam2b200_ln_proj`` (tcgen05, LayerNorm prologue); its backward uses
    ``sam2b200_ln_bwd`` / ``sam2b200_wgrad`` / the column-sum kernel;
  * the strided 3 x 3 convolutions (1 -> 4 -> 16 -> 64 -> 256 channels, < 12 % of the encoder's FLOPs) and the remaining
    1 x 1 projections go to cuDNN / cuBLAS through torch (fp32 convolutions, bf16 GEMMs with fp32 accumulation) -- the part
    of this row that is NOT hand-written yet.
CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from .memory_attention import get_clones

BF16, F32 = torch.bfloat16, torch.float32


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class _LnGeluFn(torch.autograd.Function):
    """y = act(LayerNorm_C(x) * w + b) on channels-last pixels x [..., C] fp32 (C = 4 | 16 | 64 | 256)."""

    @staticmethod
    def forward(ctx, x, w, b, eps: float, act: bool):
        if not x.is_cuda:
            raise _lib.Sam2B200Error("MemoryEncoder (B200 path) needs CUDA tensors: no CPU fallback")
        x = x.contiguous().float()
        c = x.shape[-1]
        p = x.numel() // c
        y = torch.empty_like(x)
        wf, bf = w.detach().float().contiguous(), b.detach().float().contiguous()
        rc = _lib.load().sam2b200_ln_gelu_fwd(x.data_ptr(), wf.data_ptr(), bf.data_ptr(), y.data_ptr(), p, c, float(eps), int(act),
                                              _stream(x.device))
        _lib.check(rc, "sam2b200_ln_gelu_fwd")
        ctx.save_for_backward(x, wf, bf)
        ctx.cfg = (float(eps), int(act))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wf, bf = ctx.saved_tensors
        eps, act = ctx.cfg
        c = x.shape[-1]
        p = x.numel() // c
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dw = torch.zeros(c, dtype=F32, device=x.device)
        db = torch.zeros(c, dtype=F32, device=x.device)
        rc = _lib.load().sam2b200_ln_gelu_bwd(dy.data_ptr(), x.data_ptr(), wf.data_ptr(), bf.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                              db.data_ptr(), p, c, eps, act, _stream(x.device))
        _lib.check(rc, "sam2b200_ln_gelu_bwd")
        return dx, dw, db, None, None


class _DwConv7Fn(torch.autograd.Function):
    """Depth-wise 7 x 7 convolution, padding 3, on channels-last x [B, H, W, C] fp32; w [C, 1, 7, 7], bias [C]."""

    @staticmethod
    def forward(ctx, x, w, bias):
        if not x.is_cuda:
            raise _lib.Sam2B200Error("MemoryEncoder (B200 path) needs CUDA tensors: no CPU fallback")
        x = x.contiguous().float()
        b, h, wd, c = x.shape
        wf = w.detach().float().contiguous().view(c, 49)
        bf = bias.detach().float().contiguous() if bias is not None else None
        y = torch.empty_like(x)
        rc = _lib.load().sam2b200_dwconv7(x.data_ptr(), wf.data_ptr(), bf.data_ptr() if bf is not None else None, y.data_ptr(), b, h, wd, c, 0,
                                          _stream(x.device))
        _lib.check(rc, "sam2b200_dwconv7")
        ctx.save_for_backward(x, wf)
        ctx.has_bias = bias is not None
        ctx.wshape = tuple(w.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wf = ctx.saved_tensors
        b, h, wd, c = x.shape
        lib = _lib.load()
        dy = dy.contiguous().float()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.check(lib.sam2b200_dwconv7(dy.data_ptr(), wf.data_ptr(), None, dx.data_ptr(), b, h, wd, c, 1, _stream(x.device)), "sam2b200_dwconv7")
        dw = torch.zeros(c, 49, dtype=F32, device=x.device)
        db = torch.zeros(c, dtype=F32, device=x.device)
        ws = torch.empty(max(lib.sam2b200_dwconv7_bwd_w_workspace_bytes(b, h, c) // 4, 1), dtype=F32, device=x.device)
        _lib.check(lib.sam2b200_dwconv7_bwd_w(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), b, h, wd, c,
                                              _stream(x.device)), "sam2b200_dwconv7_bwd_w")
        return dx, dw.view(ctx.wshape), (db if ctx.has_bias else None)


class _LnLinearFn(torch.autograd.Function):
    """pre [R, Nout] bf16 = LayerNorm_256(u) @ W^T + bias in ONE tcgen05 kernel (sam2b200_ln_proj); u [R, 256] fp32."""

    @staticmethod
    def forward(ctx, u, gamma, beta, weight, bias, eps: float):
        from .. import fused_stack as fs
        u = u.contiguous().float()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        w16 = weight.detach().to(BF16).contiguous()
        bias16 = bias.detach().to(BF16).contiguous()
        (pre,), y, _, mean, rstd = fs.ln_proj(u, None, g32, b32, w16, bias16, 1, out_width=w16.shape[0], eps=eps)
        ctx.save_for_backward(u, g32, y, mean, rstd, w16)
        return pre

    @staticmethod
    def backward(ctx, dpre):
        from .. import fused_stack as fs
        u, g32, y, mean, rstd, w16 = ctx.saved_tensors
        dev = u.device
        dpre = dpre.to(BF16).contiguous()
        nout = w16.shape[0]
        dw = fs.wgrad_(torch.zeros((nout, 256), dtype=F32, device=dev), dpre, y)          # dW = dpre^T y
        db = torch.zeros(nout, dtype=F32, device=dev)
        fs.bias_grad_(db, dpre)
        dy = torch.mm(dpre, w16)                                                           # [R, 256] bf16
        dgamma = torch.zeros(256, dtype=F32, device=dev)
        dbeta = torch.zeros(256, dtype=F32, device=dev)
        du = fs.ln_bwd(dy, u, mean, rstd, g32, None, dgamma, dbeta)
        return du, dgamma, dbeta, dw, db, None


class LayerNorm2d(nn.Module):  # sam2_utils.py:141-153
    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:     # [B, C, H, W] -> [B, C, H, W]
        y = _LnGeluFn.apply(x.permute(0, 2, 3, 1), self.weight, self.bias, self.eps, False)
        return y.permute(0, 3, 1, 2)


class DropPath(nn.Module):  # sam2_utils.py:92-107
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep_prob = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
        if keep_prob > 0.0 and self.scale_by_keep:
            random_tensor.div_(keep_prob)
        return x * random_tensor


class MaskDownSampler(nn.Module):
    """memory_encoder.py:17-59: ``log_stride(total_stride)`` x [Conv2d(stride) -> LayerNorm2d -> GELU], then a 1 x 1 projection."""

    def __init__(self, embed_dim=256, kernel_size=4, stride=4, padding=0, total_stride=16, activation=nn.GELU):
        super().__init__()
        num_layers = int(math.log2(total_stride) // math.log2(stride))
        assert stride ** num_layers == total_stride
        self.encoder = nn.Sequential()
        mask_in_chans, mask_out_chans = 1, 1
        for _ in range(num_layers):
            mask_out_chans = mask_in_chans * (stride ** 2)
            self.encoder.append(nn.Conv2d(mask_in_chans, mask_out_chans, kernel_size=kernel_size, stride=stride, padding=padding))
            self.encoder.append(LayerNorm2d(mask_out_chans))
            self.encoder.append(activation())
            mask_in_chans = mask_out_chans
        self.encoder.append(nn.Conv2d(mask_out_chans, embed_dim, kernel_size=1))

    def forward(self, x):
        """[B, 1, S, S] -> [B, embed_dim, S / total_stride, S / total_stride]; every LayerNorm2d + GELU pair is one fused kernel."""
        mods = list(self.encoder)
        x = x.float().contiguous(memory_format=torch.channels_last)
        i = 0
        while i < len(mods):
            m = mods[i]
            if (isinstance(m, nn.Conv2d) and i + 2 < len(mods) and isinstance(mods[i + 1], LayerNorm2d) and isinstance(mods[i + 2], nn.GELU)
                    and m.out_channels in (4, 16, 64, 256)):
                x = F.conv2d(x, m.weight, m.bias, m.stride, m.padding)                      # cuDNN, channels-last fp32
                ln = mods[i + 1]
                x = _LnGeluFn.apply(x.permute(0, 2, 3, 1), ln.weight, ln.bias, ln.eps, True).permute(0, 3, 1, 2)
                i += 3
            else:
                x = m(x)
                i += 1
        return x


class CXBlock(nn.Module):
    """memory_encoder.py:62-110 (ConvNeXt block): dwconv -> LayerNorm -> Linear(4 dim) -> GELU -> Linear -> gamma -> residual."""

    def __init__(self, dim, kernel_size=7, padding=3, drop_path=0.0, layer_scale_init_value=1e-6, use_dwconv=True):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=kernel_size, padding=padding, groups=dim if use_dwconv else 1)
        self.norm = LayerNorm2d(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = (nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True)
                      if layer_scale_init_value > 0 else None)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def _fast(self) -> bool:
        d = self.dwconv
        return (d.groups == d.in_channels == 256 and d.kernel_size == (7, 7) and d.padding == (3, 3) and d.stride == (1, 1)
                and self.pwconv1.in_features == 256 and self.pwconv1.out_features % 256 == 0 and self.pwconv1.out_features <= 2048)

    def forward_cl(self, x):
        """x: [B, H, W, 256] fp32 channels-last -> same."""
        if not self._fast():
            raise _lib.Sam2B200Error("CXBlock (B200 path): depth-wise 7 x 7 / dim 256 only (configs/sam2/sam2.1_hiera_t.yaml:78-84)")
        b, h, w, c = x.shape
        u = _DwConv7Fn.apply(x, self.dwconv.weight, self.dwconv.bias)
        pre = _LnLinearFn.apply(u.view(b * h * w, c), self.norm.weight, self.norm.bias, self.pwconv1.weight, self.pwconv1.bias, self.norm.eps)
        hid = self.act(pre)                                                                   # bf16
        y = F.linear(hid, self.pwconv2.weight.to(BF16), self.pwconv2.bias.to(BF16)).float().view(b, h, w, c)
        if self.gamma is not None:
            y = self.gamma * y
        return x + self.drop_path(y)

    def forward(self, x):      # [B, C, H, W] like the reference
        return self.forward_cl(x.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)


class Fuser(nn.Module):  # memory_encoder.py:113-131
    def __init__(self, layer, num_layers, dim=None, input_projection=False):
        super().__init__()
        self.proj = nn.Identity()
        self.layers = get_clones(layer, num_layers)
        if input_projection:
            assert dim is not None
            self.proj = nn.Conv2d(dim, dim, kernel_size=1)

    def forward(self, x):      # [B, C, H, W]
        x = self.proj(x)
        x = x.permute(0, 2, 3, 1).contiguous().float()
        for layer in self.layers:
            x = layer.forward_cl(x) if hasattr(layer, "forward_cl") else layer(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        return x.permute(0, 3, 1, 2)


class PositionEmbeddingSine(nn.Module):
    """position_encoding.py:16-130 (forward / _pe only: the image-shaped sine encoding, cached per feature size)."""

    def __init__(self, num_pos_feats, temperature: int = 10000, normalize: bool = True, scale: Optional[float] = None, **_unused):
        super().__init__()
        assert num_pos_feats % 2 == 0, "Expecting even model width"
        self.num_pos_feats = num_pos_feats // 2
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.scale = 2 * math.pi if scale is None else scale
        self.cache = {}

    @torch.no_grad()
    def _pe(self, B, device, *cache_key):
        H, W = cache_key
        if cache_key in self.cache:
            return self.cache[cache_key].to(device)[None].repeat(B, 1, 1, 1)
        y_embed = torch.arange(1, H + 1, dtype=torch.float32, device=device).view(1, -1, 1).repeat(B, 1, W)
        x_embed = torch.arange(1, W + 1, dtype=torch.float32, device=device).view(1, 1, -1).repeat(B, H, 1)
        if self.normalize:
            eps = 1e-6
            y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
            x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos = torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)
        self.cache[cache_key] = pos[0]
        return pos

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        return self._pe(x.shape[0], x.device, x.shape[-2], x.shape[-1])


class MemoryEncoder(nn.Module):
    """memory_encoder.py:134-181: masks -> MaskDownSampler, + pix_feat_proj(pix_feat), Fuser (2 CXBlocks), out_proj (256 -> 64),
    sine position encoding.  Returns ``{"vision_features": [B, out_dim, H, W], "vision_pos_enc": [pos]}`` like the reference."""

    def __init__(self, out_dim, mask_downsampler, fuser, position_encoding, in_dim=256):
        super().__init__()
        self.mask_downsampler = mask_downsampler
        self.pix_feat_proj = nn.Conv2d(in_dim, in_dim, kernel_size=1)
        self.fuser = fuser
        self.position_encoding = position_encoding
        self.out_proj = nn.Identity()
        if out_dim != in_dim:
            self.out_proj = nn.Conv2d(in_dim, out_dim, kernel_size=1)

    @staticmethod
    def _conv1x1(x_cl, conv: nn.Conv2d):
        """1 x 1 convolution on channels-last tokens [B, H, W, Cin] as one bf16 GEMM with fp32 accumulation."""
        w = conv.weight.view(conv.out_channels, conv.in_channels)
        return F.linear(x_cl.to(BF16), w.to(BF16), conv.bias.to(BF16) if conv.bias is not None else None).float()

    def forward(self, pix_feat: torch.Tensor, masks: torch.Tensor, skip_mask_sigmoid: bool = False) -> dict:
        if not masks.is_cuda:
            raise _lib.Sam2B200Error("MemoryEncoder (B200 path) needs CUDA tensors: no CPU fallback")
        _lib.load()
        if not skip_mask_sigmoid:
            masks = torch.sigmoid(masks)                      # memory_encoder.py:160-161
        m = self.mask_downsampler(masks)                      # [B, 256, H, W] (channels-last memory)
        pix_feat = pix_feat.to(m.device)
        x = self._conv1x1(pix_feat.permute(0, 2, 3, 1), self.pix_feat_proj) + m.permute(0, 2, 3, 1)      # :169-170, channels-last
        if isinstance(self.fuser, Fuser) and isinstance(self.fuser.proj, nn.Identity):
            for layer in self.fuser.layers:
                x = layer.forward_cl(x.contiguous()) if hasattr(layer, "forward_cl") else layer(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        else:
            x = self.fuser(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        if isinstance(self.out_proj, nn.Conv2d):
            x = self._conv1x1(x, self.out_proj)
        x = x.permute(0, 3, 1, 2)                             # [B, out_dim, H, W] view of the token-major buffer
        pos = self.position_encoding(x).to(x.dtype)
        return {"vision_features": x, "vision_pos_enc": [pos]}


def build_memory_encoder(out_dim: int = 64) -> MemoryEncoder:
    """The memory encoder of configs/sam2/sam2.1_hiera_t.yaml:62-85 (same for every SAM2.1 size)."""
    return MemoryEncoder(
        out_dim=out_dim,
        position_encoding=PositionEmbeddingSine(num_pos_feats=64, normalize=True, scale=None, temperature=10000),
        mask_downsampler=MaskDownSampler(kernel_size=3, stride=2, padding=1),
        fuser=Fuser(layer=CXBlock(dim=256, kernel_size=7, padding=3, layer_scale_init_value=1e-6, use_dwconv=True), num_layers=2))
