"""CPU oracle for the SAM2 memory encoder (TEST INFRASTRUCTURE ONLY).

A functional restatement of ``sam2_video/model/modeling/memory_encoder.py`` of the reference (MaskDownSampler :17-59,
CXBlock :62-110, Fuser :113-131, MemoryEncoder :134-181), LayerNorm2d (sam2_utils.py:141-153) and PositionEmbeddingSine
(position_encoding.py:90-130) on a plain ``{name: tensor}`` state dict with the reference's keys -- no nn.Module, explicit
formulas for the normalisations / GELU / layer scale; convolutions through ``F.conv2d``.  Differentiable (autograd), any
dtype / device.  Only ``tests/`` and ``__graft_entry__.smoke()`` may import this module.

Parity status: PINNED against the unmodified reference classes executed through ``oracle/ref_shim.py``
(``oracle/make_golden.py::golden_memory_encoder`` -> tests/golden/memenc_*.npz, checked by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def layer_norm_2d(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """sam2_utils.py:148-153: normalise over the channel dim of [B, C, H, W]."""
    u = x.mean(1, keepdim=True)
    s = ((x - u) ** 2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def gelu(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))      # nn.GELU() default (exact)


def mask_downsampler(p: Dict[str, Tensor], x: Tensor, prefix: str = "mask_downsampler.encoder.", stride: int = 2, padding: int = 1) -> Tensor:
    """memory_encoder.py:38-59 with kernel_size 3 / stride 2 / padding 1 (configs/sam2/sam2.1_hiera_t.yaml:71-75): 4 x
    [conv -> LayerNorm2d -> GELU] then the 1 x 1 projection (encoder indices 0..11 and 12)."""
    i = 0
    while f"{prefix}{i + 1}.weight" in p and p[f"{prefix}{i + 1}.weight"].dim() == 1:
        x = F.conv2d(x, p[f"{prefix}{i}.weight"], p[f"{prefix}{i}.bias"], stride=stride, padding=padding)
        x = gelu(layer_norm_2d(x, p[f"{prefix}{i + 1}.weight"], p[f"{prefix}{i + 1}.bias"]))
        i += 3
    return F.conv2d(x, p[f"{prefix}{i}.weight"], p[f"{prefix}{i}.bias"])


def cx_block(p: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """memory_encoder.py:97-110: x + gamma * pwconv2(GELU(pwconv1(LayerNorm(dwconv7x7(x))))) (channels-last linears)."""
    c = x.shape[1]
    h = F.conv2d(x, p[prefix + "dwconv.weight"], p[prefix + "dwconv.bias"], padding=3, groups=c)
    h = layer_norm_2d(h, p[prefix + "norm.weight"], p[prefix + "norm.bias"])
    h = h.permute(0, 2, 3, 1)
    h = h @ p[prefix + "pwconv1.weight"].t() + p[prefix + "pwconv1.bias"]
    h = gelu(h)
    h = h @ p[prefix + "pwconv2.weight"].t() + p[prefix + "pwconv2.bias"]
    h = p[prefix + "gamma"] * h
    return x + h.permute(0, 3, 1, 2)


def position_embedding_sine(b: int, h: int, w: int, num_pos_feats: int = 64, temperature: float = 10000.0, dtype=torch.float32) -> Tensor:
    """position_encoding.py:90-124 (normalize=True, scale 2 pi): [B, num_pos_feats, H, W]."""
    half = num_pos_feats // 2
    y_embed = torch.arange(1, h + 1, dtype=torch.float32).view(1, -1, 1).repeat(b, 1, w)
    x_embed = torch.arange(1, w + 1, dtype=torch.float32).view(1, 1, -1).repeat(b, h, 1)
    y_embed = y_embed / (y_embed[:, -1:, :] + 1e-6) * (2 * math.pi)
    x_embed = x_embed / (x_embed[:, :, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(half, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / half)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2).to(dtype)


def memory_encoder(p: Dict[str, Tensor], pix_feat: Tensor, masks: Tensor, skip_mask_sigmoid: bool = False, num_layers: int = 2):
    """MemoryEncoder.forward (memory_encoder.py:154-181) -> (vision_features [B, 64, H, W], vision_pos_enc [B, 64, H, W])."""
    if not skip_mask_sigmoid:
        masks = torch.sigmoid(masks)
    m = mask_downsampler(p, masks)
    x = F.conv2d(pix_feat, p["pix_feat_proj.weight"], p["pix_feat_proj.bias"]) + m
    for i in range(num_layers):
        x = cx_block(p, f"fuser.layers.{i}.", x)
    x = F.conv2d(x, p["out_proj.weight"], p["out_proj.bias"])
    pos = position_embedding_sine(x.shape[0], x.shape[2], x.shape[3], 64, dtype=x.dtype).to(x.device)
    return x, pos


def reference_init_state(seed: int = 0) -> Dict[str, Tensor]:
    """The reference's own random initialisation (nn.Conv2d / nn.Linear defaults under torch.manual_seed(seed), modules created
    in the order the yaml lists them: mask_downsampler, fuser.layer (cloned twice), then pix_feat_proj, out_proj) with the
    layer-scale vectors gamma moved from their 1e-6 init to 0.5 +- 0.1 (so that the blocks matter in a parity test)."""
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    cin = 1
    for i in range(4):
        conv = torch.nn.Conv2d(cin, cin * 4, kernel_size=3, stride=2, padding=1)
        sd[f"mask_downsampler.encoder.{3 * i}.weight"] = conv.weight.detach().clone()
        sd[f"mask_downsampler.encoder.{3 * i}.bias"] = conv.bias.detach().clone()
        sd[f"mask_downsampler.encoder.{3 * i + 1}.weight"] = torch.ones(cin * 4)
        sd[f"mask_downsampler.encoder.{3 * i + 1}.bias"] = torch.zeros(cin * 4)
        cin *= 4
    conv = torch.nn.Conv2d(cin, 256, kernel_size=1)
    sd["mask_downsampler.encoder.12.weight"], sd["mask_downsampler.encoder.12.bias"] = conv.weight.detach().clone(), conv.bias.detach().clone()
    dw = torch.nn.Conv2d(256, 256, kernel_size=7, padding=3, groups=256)
    pw1, pw2 = torch.nn.Linear(256, 1024), torch.nn.Linear(1024, 256)
    pix = torch.nn.Conv2d(256, 256, kernel_size=1)
    out = torch.nn.Conv2d(256, 64, kernel_size=1)
    sd["pix_feat_proj.weight"], sd["pix_feat_proj.bias"] = pix.weight.detach().clone(), pix.bias.detach().clone()
    for i in range(2):
        pre = f"fuser.layers.{i}."
        sd[pre + "gamma"] = 0.5 + 0.1 * torch.sin(torch.arange(256, dtype=torch.float32) * (0.37 + i))
        sd[pre + "dwconv.weight"], sd[pre + "dwconv.bias"] = dw.weight.detach().clone(), dw.bias.detach().clone()
        sd[pre + "norm.weight"], sd[pre + "norm.bias"] = torch.ones(256), torch.zeros(256)
        sd[pre + "pwconv1.weight"], sd[pre + "pwconv1.bias"] = pw1.weight.detach().clone(), pw1.bias.detach().clone()
        sd[pre + "pwconv2.weight"], sd[pre + "pwconv2.bias"] = pw2.weight.detach().clone(), pw2.bias.detach().clone()
    sd["out_proj.weight"], sd["out_proj.bias"] = out.weight.detach().clone(), out.bias.detach().clone()
    return sd


def random_inputs(b: int, grid: int, seed: int = 77):
    """pix_feat ~ N(0, 1) [B, 256, g, g]; mask logits ~ N(0, 4^2) [B, 1, 16 g, 16 g] (SAM logits span +-10..20); upstream gradient
    ~ N(0, 1) [B, 64, g, g]."""
    g = torch.Generator().manual_seed(seed)
    return dict(pix_feat=torch.randn(b, 256, grid, grid, generator=g), masks=torch.randn(b, 1, 16 * grid, 16 * grid, generator=g) * 4.0,
                grad_out=torch.randn(b, 64, grid, grid, generator=g))
