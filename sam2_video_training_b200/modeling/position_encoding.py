"""Axial RoPE table for the B200 path.

Mirrors ``init_t_xy`` / ``compute_axial_cis`` of the reference
(sam2_video/model/modeling/position_encoding.py:185-201) but returns the rotation as a real
``[N, dim/2, 2]`` (cos, sin) fp32 table -- the layout ``sam2b200_rope_apply`` reads -- instead of a
complex64 tensor.  The rotation itself (position_encoding.py:212-239, ``apply_rotary_enc``) runs in
the CUDA library (csrc/attn.cu, rope_kernel), in fp32 registers like the reference's upcast.
"""
from __future__ import annotations

import torch


def init_t_xy(end_x: int, end_y: int):
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    t_x = (t % end_x).float()
    t_y = torch.div(t, end_x, rounding_mode="floor").float()
    return t_x, t_y


def compute_axial_cis(dim: int, end_x: int, end_y: int, theta: float = 10000.0) -> torch.Tensor:
    """-> [end_x*end_y, dim/2, 2] fp32: (cos, sin) of x*f_j for the first dim/4 pairs and of y*f_j for
    the last dim/4, f_j = theta^(-4j/dim).  Angles are formed in fp32 like the reference."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: (dim // 4)].float() / dim))
    t_x, t_y = init_t_xy(int(end_x), int(end_y))
    ang = torch.cat([torch.outer(t_x, freqs), torch.outer(t_y, freqs)], dim=-1)
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
