"""``Attention`` / ``RoPEAttention`` with the reference's constructor arguments, attributes and
state_dict keys (sam2_video/model/modeling/sam/transformer.py:190-311), computing on the sm_100a
kernels of libsam2b200.so.

Numerics: parameters stay fp32 (checkpoint compatible both ways); projections run in bf16 on the
tensor cores (cuBLAS through ``F.linear``); rotation + softmax(QK^T/16)V + their backward run in
the hand-written tcgen05 kernels (bf16 operands, fp32 accumulation).  One head of width 256 only --
the only configuration SAM2's memory attention uses (configs/sam2/sam2.1_hiera_t.yaml:40-60).
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from ... import _lib
from ...ops import RopeAttentionFn
from ..position_encoding import compute_axial_cis

class Attention(nn.Module):
    """transformer.py:190-248.  Plain (un-rotated) attention through the same fused kernel."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1, dropout: float = 0.0,
                 kv_in_dim: int = None) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.kv_in_dim = kv_in_dim if kv_in_dim is not None else embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        if self.internal_dim // num_heads != 256 or num_heads != 1:
            raise _lib.Sam2B200Error(
                "the B200 attention kernels are built for SAM2 memory attention: 1 head of width 256 "
                f"(got num_heads={num_heads}, head_dim={self.internal_dim // max(num_heads, 1)})")
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(self.kv_in_dim, self.internal_dim)
        self.v_proj = nn.Linear(self.kv_in_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)
        self.dropout_p = dropout
        self.attn_nsplit = 0  # 0 = library heuristic

    @staticmethod
    def _lin(x: Tensor, lin: nn.Linear) -> Tensor:
        return F.linear(x.to(torch.bfloat16), lin.weight.to(torch.bfloat16), lin.bias.to(torch.bfloat16))

    def _identity_table(self, device) -> Tensor:
        t = getattr(self, "_id_table", None)
        if t is None or t.device != device:
            t = torch.zeros(1, 128, 2, dtype=torch.float32, device=device)
            t[..., 0] = 1.0
            self._id_table = t
        return t

    def forward(self, q: Tensor, k: Tensor, v: Tensor) -> Tensor:
        drop_p = self.dropout_p if self.training else 0.0    # transformer.py:243: dropout_p = self.dropout_p if self.training else 0.0
        q = self._lin(q, self.q_proj)
        k = self._lin(k, self.k_proj)
        v = self._lin(v, self.v_proj)
        out = RopeAttentionFn.apply(q, k, v, self._identity_table(q.device), k.shape[1], self.attn_nsplit, drop_p)
        return self._lin(out, self.out_proj)


class RoPEAttention(Attention):
    """transformer.py:251-311: attention with axial rotary position encoding."""

    def __init__(self, *args, rope_theta=10000.0, rope_k_repeat=False, feat_sizes=(64, 64), **kwargs):
        super().__init__(*args, **kwargs)
        self.compute_cis = partial(compute_axial_cis, dim=self.internal_dim // self.num_heads, theta=rope_theta)
        # plain attribute like the reference's freqs_cis (NOT in the state_dict, transformer.py:269-272)
        self.freqs_cis = self.compute_cis(end_x=feat_sizes[0], end_y=feat_sizes[1])
        self.rope_k_repeat = rope_k_repeat
        self.rope_theta = float(rope_theta)        # kept for integrate.use_b200_attention / _fused_eligible

    def _table(self, n_tokens: int, device) -> Tensor:
        if self.freqs_cis.shape[0] != n_tokens:  # transformer.py:289-292
            w = math.sqrt(n_tokens)
            if int(w) * int(w) != n_tokens:
                raise ValueError(f"axial RoPE needs a square token grid, got {n_tokens} tokens")
            self.freqs_cis = self.compute_cis(end_x=int(w), end_y=int(w))
        if self.freqs_cis.device != device:
            self.freqs_cis = self.freqs_cis.to(device)
        return self.freqs_cis

    def forward(self, q: Tensor, k: Tensor, v: Tensor, num_k_exclude_rope: int = 0) -> Tensor:
        drop_p = self.dropout_p if self.training else 0.0    # transformer.py:304
        q = self._lin(q, self.q_proj)
        k = self._lin(k, self.k_proj)
        v = self._lin(v, self.v_proj)
        if q.shape[-2] != k.shape[-2]:
            assert self.rope_k_repeat  # transformer.py:293-294
        table = self._table(q.shape[-2], q.device)
        num_k_rope = k.shape[-2] - num_k_exclude_rope
        if num_k_rope % q.shape[-2] != 0:
            raise ValueError("rotated key count must be a multiple of the query count (position_encoding.py:230)")
        out = RopeAttentionFn.apply(q, k, v, table, num_k_exclude_rope, self.attn_nsplit, drop_p)
        return self._lin(out, self.out_proj)
