"""``MemoryAttentionLayer`` / ``MemoryAttention`` with the reference's constructor arguments,
attribute names, state_dict keys (106 tensors, 5 922 304 parameters) and forward signatures
(sam2_video/model/modeling/memory_attention.py:17-169), running on the B200 kernels.

The residual stream is kept in fp32; LayerNorm runs in fp32 and hands bf16 to the projections;
attention runs in libsam2b200.so.  Requires CUDA tensors -- there is no CPU fallback.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .. import _lib
from .sam.transformer import RoPEAttention


def get_activation_fn(activation):  # sam2_utils.py:77-85
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu, not {activation}.")


def get_clones(module, N):  # sam2_utils.py:88-89
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class MemoryAttentionLayer(nn.Module):
    def __init__(self, activation: str, cross_attention: nn.Module, d_model: int, dim_feedforward: int,
                 dropout: float, pos_enc_at_attn: bool, pos_enc_at_cross_attn_keys: bool,
                 pos_enc_at_cross_attn_queries: bool, self_attention: nn.Module):
        super().__init__()
        self.d_model = d_model
        self.dim_feedforward = dim_feedforward
        self.dropout_value = dropout
        self.self_attn = self_attention
        self.cross_attn_image = cross_attention
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.activation_str = activation
        self.activation = get_activation_fn(activation)
        self.pos_enc_at_attn = pos_enc_at_attn
        self.pos_enc_at_cross_attn_queries = pos_enc_at_cross_attn_queries
        self.pos_enc_at_cross_attn_keys = pos_enc_at_cross_attn_keys

    @staticmethod
    def _ln(x: Tensor, ln: nn.LayerNorm) -> Tensor:
        return F.layer_norm(x.float(), ln.normalized_shape, ln.weight, ln.bias, ln.eps)

    def _forward_sa(self, tgt, query_pos):  # memory_attention.py:58-64
        tgt2 = self._ln(tgt, self.norm1)
        q = k = tgt2 + query_pos if self.pos_enc_at_attn else tgt2
        tgt2 = self.self_attn(q, k, v=tgt2)
        return tgt + self.dropout1(tgt2.float())

    def _forward_ca(self, tgt, memory, query_pos, pos, num_k_exclude_rope=0):  # memory_attention.py:66-81
        kwds = {}
        if num_k_exclude_rope > 0:
            assert isinstance(self.cross_attn_image, RoPEAttention)
            kwds = {"num_k_exclude_rope": num_k_exclude_rope}
        tgt2 = self._ln(tgt, self.norm2)
        tgt2 = self.cross_attn_image(
            q=tgt2 + query_pos if self.pos_enc_at_cross_attn_queries else tgt2,
            k=memory + pos if self.pos_enc_at_cross_attn_keys else memory,
            v=memory, **kwds)
        return tgt + self.dropout2(tgt2.float())

    def forward(self, tgt, memory, pos: Optional[Tensor] = None, query_pos: Optional[Tensor] = None,
                num_k_exclude_rope: int = 0) -> torch.Tensor:
        tgt = self._forward_sa(tgt, query_pos)
        tgt = self._forward_ca(tgt, memory, query_pos, pos, num_k_exclude_rope)
        tgt2 = self._ln(tgt, self.norm3).to(torch.bfloat16)
        h = self.activation(F.linear(tgt2, self.linear1.weight.to(torch.bfloat16), self.linear1.bias.to(torch.bfloat16)))
        tgt2 = F.linear(self.dropout(h), self.linear2.weight.to(torch.bfloat16), self.linear2.bias.to(torch.bfloat16))
        return tgt + self.dropout3(tgt2.float())


class MemoryAttention(nn.Module):
    def __init__(self, d_model: int, pos_enc_at_input: bool, layer: nn.Module, num_layers: int,
                 batch_first: bool = True):
        super().__init__()
        self.d_model = d_model
        self.layers = get_clones(layer, num_layers)
        self.num_layers = num_layers
        self.norm = nn.LayerNorm(d_model)
        self.pos_enc_at_input = pos_enc_at_input
        self.batch_first = batch_first
        # True: one hand-scheduled autograd.Function for the whole stack (fused_stack.py) whenever the
        # stack has the shipped SAM2 configuration; False: compose the per-module path below.
        self.use_fused_stack = True
        self.attn_nsplit = 0

    def _fused_eligible(self) -> bool:
        for layer in self.layers:
            if not isinstance(layer, MemoryAttentionLayer):
                return False
            sa, ca = layer.self_attn, layer.cross_attn_image
            if not (type(sa) is RoPEAttention and type(ca) is RoPEAttention):
                return False
            if (layer.pos_enc_at_attn or not layer.pos_enc_at_cross_attn_keys or layer.pos_enc_at_cross_attn_queries
                    or layer.activation_str != "relu" or not ca.rope_k_repeat or ca.kv_in_dim != 64
                    or layer.d_model != 256 or not self.batch_first):
                return False
            # hyper-parameters the kernels hard-code (csrc/glue.cu: eps 1e-5, width 256; csrc/mlp.cu: hidden width 2048;
            # one RoPE table for both attentions): anything else takes the composed path instead of wrong numerics
            if (layer.dim_feedforward != 2048 or layer.linear1.out_features != 2048
                    or any(abs(ln.eps - 1e-5) > 1e-12 for ln in (layer.norm1, layer.norm2, layer.norm3, self.norm))
                    or getattr(sa, "rope_theta", 10000.0) != getattr(ca, "rope_theta", 10000.0)
                    or sa.num_heads != 1 or ca.num_heads != 1 or sa.internal_dim != 256 or ca.internal_dim != 256):
                return False
            l0 = self.layers[0]   # one set of dropout rates for the stack (get_clones copies the layer)
            if (layer.dropout_value, sa.dropout_p, ca.dropout_p) != (l0.dropout_value, l0.self_attn.dropout_p,
                                                                      l0.cross_attn_image.dropout_p):
                return False
        return True

    def grad_bucket_order(self):
        """Layout ddp.GradBucket should use: per layer the q/k/v self-attention weights back to back, then their
        biases back to back (one [768, 256] weight-gradient GEMM / one bias reduction writes all three), then the rest."""
        order = []
        for layer in self.layers:
            sa = getattr(layer, "self_attn", None)
            if sa is not None and all(hasattr(sa, k) for k in ("q_proj", "k_proj", "v_proj")):
                order += [sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight, sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias]
        return order

    def _grad_anchor(self, device):
        a = getattr(self, "_sam2b200_anchor", None)
        if a is None or a.device != device:
            a = torch.zeros(1, device=device, requires_grad=True)
            self._sam2b200_anchor = a
        return a

    def _forward_fused(self, curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens, anchor=None, packed=None, raw=False,
                       direct=None):
        """packed = (tpos_rows, obj_pos, n_slots, hw) of a memory_bank.PackedBank: `memory` / `memory_pos` are then its memk /
        memv tensors ([B, M, 64] bf16)."""
        from ..fused_stack import MemoryAttentionStackFn
        n = curr.shape[0]
        m = memory.shape[1] if packed else memory.shape[0]
        ca0 = self.layers[0].cross_attn_image
        if (m - num_obj_ptr_tokens) % n != 0:
            raise ValueError("rotated key count must be a multiple of the query count (position_encoding.py:230)")
        table = ca0._table(n, curr.device)
        from ..fused_stack import direct_grads_possible
        params = [p for _, p in self.named_parameters()]
        bucket = getattr(self, "_sam2b200_grad_bucket", None)
        # With a GradBucket attached (ddp.attach_grad_bucket) the backward accumulates parameter gradients into the
        # bucket itself; autograd then only sees detached aliases of the parameters (no 106 AccumulateGrad nodes).
        if direct is None:     # (graphs.py decides once per captured signature and passes its decision in: it builds the call
            direct = direct_grads_possible(bucket, params)      # under no_grad, where this test would say no)
        # train-mode dropout: one device seed per call; the kernels derive every mask of the stack from it
        dropout = None
        l0 = self.layers[0]
        if self.training and max(l0.dropout_value, l0.self_attn.dropout_p, l0.cross_attn_image.dropout_p) > 0:
            dropout = dict(seed=torch.empty(1, dtype=torch.int64, device=curr.device).random_(),
                           p_res=float(l0.dropout_value), p_sa=float(l0.self_attn.dropout_p),
                           p_ca=float(l0.cross_attn_image.dropout_p))
            if getattr(self, "_sam2b200_fixed_seed", None) is not None:   # tests: reproduce the masks
                dropout["seed"].fill_(int(self._sam2b200_fixed_seed))
        meta = dict(num_layers=self.num_layers, num_k_exclude_rope=int(num_obj_ptr_tokens), table=table,
                    pos_enc_at_input=bool(self.pos_enc_at_input), nsplit=int(self.attn_nsplit),
                    bucket=bucket, direct=direct, master_params=params, dropout=dropout, packed=packed is not None,
                    bank_slots=packed[2] if packed is not None else 0, bank_hw=packed[3] if packed is not None else 0)
        b_tpos, b_obj = (packed[0], packed[1]) if packed is not None else (None, None)
        if b_tpos is not None and b_tpos.numel() == 0:
            b_tpos = None
        if b_obj is not None and b_obj.numel() == 0:
            b_obj = None
        if direct:
            anchor = anchor if anchor is not None else self._grad_anchor(curr.device)
            args = (meta, curr, curr_pos, memory, memory_pos, b_tpos, b_obj, *[p.detach() for p in params], anchor)
        else:
            args = (meta, curr, curr_pos, memory, memory_pos, b_tpos, b_obj, *params)
        if raw:         # graphs.GraphedMemoryAttention calls the function's forward / backward itself (no autograd in a capture)
            return args
        return MemoryAttentionStackFn.apply(*args)

    def forward(self, curr: torch.Tensor, memory: torch.Tensor, curr_pos: Optional[Tensor] = None,
                memory_pos: Optional[Tensor] = None, num_obj_ptr_tokens: int = 0):
        if isinstance(curr, list):  # memory_attention.py:127-133
            assert isinstance(curr_pos, list)
            assert len(curr) == len(curr_pos) == 1
            curr, curr_pos = curr[0], curr_pos[0]
        assert curr.shape[1] == memory.shape[1], "Batch size must be the same for curr and memory"
        if not curr.is_cuda:
            raise _lib.Sam2B200Error("MemoryAttention (B200 path) needs CUDA tensors: no CPU fallback")
        _lib.load()  # fail loudly before any compute if the CUDA library is missing
        from ..memory_bank import PackedBank
        if isinstance(memory, PackedBank):
            # the bank already is in the kernels' layout (memory_bank.assemble_memory_packed): memory_pos is ignored
            if not (self.use_fused_stack and self._fused_eligible()):
                raise _lib.Sam2B200Error("a PackedBank needs the fused stack (shipped SAM2 configuration)")
            return self._forward_fused(curr, memory.memk, curr_pos, memory.memv, memory.num_obj_ptr_tokens,
                                       packed=(memory.tpos_rows, memory.obj_pos, memory.n_slots, memory.hw))
        if self.use_fused_stack and self._fused_eligible():
            return self._forward_fused(curr, memory, curr_pos, memory_pos, num_obj_ptr_tokens)
        output = curr.float()
        if self.pos_enc_at_input and curr_pos is not None:
            output = output + 0.1 * curr_pos
        if self.batch_first:
            output = output.transpose(0, 1)
            curr_pos = curr_pos.transpose(0, 1)
            memory = memory.transpose(0, 1)
            memory_pos = memory_pos.transpose(0, 1)
        for layer in self.layers:
            kwds = {}
            if isinstance(layer.cross_attn_image, RoPEAttention):
                kwds = {"num_k_exclude_rope": num_obj_ptr_tokens}
            output = layer(tgt=output, memory=memory, pos=memory_pos, query_pos=curr_pos, **kwds)
        normed_output = F.layer_norm(output, self.norm.normalized_shape, self.norm.weight, self.norm.bias, self.norm.eps)
        if self.batch_first:
            normed_output = normed_output.transpose(0, 1)
        return normed_output


def build_memory_attention(dropout: float = 0.1, feat_sizes=(64, 64), num_layers: int = 4, sa_dropout: Optional[float] = None,
                           ca_dropout: Optional[float] = None, dim_feedforward: int = 2048, rope_theta: float = 10000.0,
                           ca_rope_theta: Optional[float] = None, pos_enc_at_input: bool = True) -> MemoryAttention:
    """The stack of configs/sam2/sam2.1_hiera_t.yaml:29-60 (same for every SAM2.1 size); the keyword arguments cover the
    hyper-parameters a differently configured reference stack may carry (integrate.use_b200_attention reads them)."""
    sa = RoPEAttention(rope_theta=rope_theta, feat_sizes=list(feat_sizes), embedding_dim=256, num_heads=1,
                       downsample_rate=1, dropout=dropout if sa_dropout is None else sa_dropout)
    ca = RoPEAttention(rope_theta=rope_theta if ca_rope_theta is None else ca_rope_theta, feat_sizes=list(feat_sizes),
                       rope_k_repeat=True, embedding_dim=256, num_heads=1, downsample_rate=1,
                       dropout=dropout if ca_dropout is None else ca_dropout, kv_in_dim=64)
    layer = MemoryAttentionLayer(activation="relu", dim_feedforward=dim_feedforward, dropout=dropout, pos_enc_at_attn=False,
                                 self_attention=sa, d_model=256, pos_enc_at_cross_attn_keys=True,
                                 pos_enc_at_cross_attn_queries=False, cross_attention=ca)
    return MemoryAttention(d_model=256, pos_enc_at_input=pos_enc_at_input, layer=layer, num_layers=num_layers)
