"""CPU oracle for the memory-attention hot path (TEST INFRASTRUCTURE ONLY).

A plain restatement -- explicit matmul / softmax / real-valued rotation, no
``scaled_dot_product_attention``, no complex tensors, no nn.Module -- of the
reference's SAM2 ``MemoryAttention`` stack.  Every function cites the reference
file:line it follows (paths relative to the reference root).  It runs in fp32 or
fp64 on the CPU and is differentiable through torch autograd so that gradient
parity can be checked as well.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
package never does (it fails loudly when its CUDA library is missing).

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified
reference modules executed in the build container through ``oracle/ref_shim.py``;
the fixtures and the script that generated them are ``tests/golden/*.npz`` and
``oracle/make_golden.py``; ``tests/test_oracle_golden.py`` re-checks them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor

# hyper-parameters of configs/sam2/sam2.1_hiera_t.yaml:29-60 (identical for all SAM2.1 sizes)
D_MODEL = 256
KV_IN_DIM = 64
DIM_FF = 2048
NUM_LAYERS = 4
ROPE_THETA = 10000.0
LN_EPS = 1e-5


def axial_rope_table(n_tokens: int, dim: int = D_MODEL, theta: float = ROPE_THETA,
                     dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """cos/sin tables ``[N, dim/2]`` of the axial rotation.

    position_encoding.py:185-201 (``init_t_xy`` + ``compute_axial_cis``): token i sits at
    x = i mod W, y = floor(i / W) with W = H = sqrt(N) (transformer.py:289-292 recomputes the
    table from sqrt of the query length); 64 frequencies f_j = theta^(-4j/dim); the first dim/4
    complex pairs rotate by x*f_j, the last dim/4 by y*f_j.  The reference builds the angles in
    fp32 (outer product of fp32 tensors) and then takes cos/sin (torch.polar) -- we do the same
    before casting to ``dtype`` so that fp64 runs see the very same table.
    """
    w = math.sqrt(n_tokens)
    if int(w) * int(w) != n_tokens:
        raise ValueError("axial RoPE needs a square token grid (transformer.py:289)")
    w = int(w)
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(n_tokens, dtype=torch.float32)
    t_x = (t % w).float()
    t_y = torch.div(t, w, rounding_mode="floor").float()
    ang = torch.cat([torch.outer(t_x, freqs), torch.outer(t_y, freqs)], dim=-1)  # [N, dim/2] fp32
    return torch.cos(ang).to(dtype), torch.sin(ang).to(dtype)


def apply_axial_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """Rotate adjacent (even, odd) pairs of the last dim.

    position_encoding.py:212-239 (``apply_rotary_enc``): ``view_as_complex`` pairs elements
    (2j, 2j+1); multiplication by cos+i*sin gives
    ``out[2j] = x[2j]*c - x[2j+1]*s``, ``out[2j+1] = x[2j]*s + x[2j+1]*c``.
    ``x``: [B, L, dim] with L a multiple of the table length (keys tile the table,
    position_encoding.py:230-237).
    """
    b, l, d = x.shape
    n = cos.shape[0]
    assert l % n == 0
    r = l // n
    if r > 1:
        cos = cos.repeat(r, 1)
        sin = sin.repeat(r, 1)
    xe = x[..., 0::2]
    xo = x[..., 1::2]
    oe = xe * cos - xo * sin
    oo = xe * sin + xo * cos
    return torch.stack([oe, oo], dim=-1).reshape(b, l, d)


def layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """nn.LayerNorm(256), eps 1e-5 (memory_attention.py:43-45,115)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def linear(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return x @ w.t() + b


def _drop(x: Tensor, keep: Optional[Tensor], p_drop: float) -> Tensor:
    """Inverted dropout with an EXPLICIT keep mask (nn.Dropout / SDPA dropout_p in train mode:
    memory_attention.py:64,81,97,99, transformer.py:304-306): x * keep / (1 - p)."""
    if keep is None:
        return x
    return x * keep.to(x.dtype) / (1.0 - p_drop)


def rope_attention(p: Dict[str, Tensor], prefix: str, q_in: Tensor, k_in: Tensor, v_in: Tensor,
                   num_k_exclude_rope: int, rope_k_repeat: bool,
                   return_parts: bool = False, keep: Optional[Tensor] = None, p_drop: float = 0.0):
    """``RoPEAttention.forward`` (transformer.py:275-311), one head of width 256.

    q/k/v projections (:277-279), head split is a no-op for one head (:282-284), rotate q and
    the first ``M - num_k_exclude_rope`` keys (:296-302), softmax(q k^T / sqrt(256)) v (:306,
    SDPA default scale, no mask, dropout 0 for parity), out projection (:308-309).
    """
    q = linear(q_in, p[prefix + "q_proj.weight"], p[prefix + "q_proj.bias"])
    k = linear(k_in, p[prefix + "k_proj.weight"], p[prefix + "k_proj.bias"])
    v = linear(v_in, p[prefix + "v_proj.weight"], p[prefix + "v_proj.bias"])
    n, m = q.shape[1], k.shape[1]
    if n != m and not rope_k_repeat:
        raise AssertionError("rope_k_repeat required when N != M (transformer.py:293-294)")
    cos, sin = (t.to(q.device) for t in axial_rope_table(n, q.shape[-1], dtype=q.dtype))
    num_k_rope = m - num_k_exclude_rope
    q = apply_axial_rope(q, cos, sin)
    if num_k_rope > 0:
        k = torch.cat([apply_axial_rope(k[:, :num_k_rope], cos, sin), k[:, num_k_rope:]], dim=1)
    s = (q @ k.transpose(1, 2)) / math.sqrt(q.shape[-1])
    a = _drop(torch.softmax(s, dim=-1), keep, p_drop)   # SDPA drops normalised probabilities (transformer.py:304-306)
    o = a @ v
    out = linear(o, p[prefix + "out_proj.weight"], p[prefix + "out_proj.bias"])
    if return_parts:
        return out, dict(q=q, k=k, v=v, o=o)
    return out


def memory_attention_layer(p: Dict[str, Tensor], prefix: str, tgt: Tensor, memory: Tensor,
                           pos: Tensor, query_pos: Tensor, num_k_exclude_rope: int,
                           masks: Optional[Dict[str, Tensor]] = None, p_drop: float = 0.0) -> Tensor:
    """``MemoryAttentionLayer.forward`` (memory_attention.py:58-99) with the shipped flags
    pos_enc_at_attn=False, pos_enc_at_cross_attn_keys=True, pos_enc_at_cross_attn_queries=False
    (configs/sam2/sam2.1_hiera_t.yaml:38,48-49), ReLU MLP, dropout 0."""
    mk = masks or {}    # train mode: keep masks "sa_prob", "ca_prob", "drop1", "drop2", "mlp", "drop3" (None = no dropout)
    t2 = layer_norm(tgt, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])
    t2 = rope_attention(p, prefix + "self_attn.", t2, t2, t2, 0, rope_k_repeat=False, keep=mk.get("sa_prob"), p_drop=p_drop)
    tgt = tgt + _drop(t2, mk.get("drop1"), p_drop)                                   # :64
    t2 = layer_norm(tgt, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"])
    t2 = rope_attention(p, prefix + "cross_attn_image.", t2, memory + pos, memory,
                        num_k_exclude_rope, rope_k_repeat=True, keep=mk.get("ca_prob"), p_drop=p_drop)
    tgt = tgt + _drop(t2, mk.get("drop2"), p_drop)                                   # :81
    t2 = layer_norm(tgt, p[prefix + "norm3.weight"], p[prefix + "norm3.bias"])
    hid = _drop(torch.relu(linear(t2, p[prefix + "linear1.weight"], p[prefix + "linear1.bias"])), mk.get("mlp"), p_drop)   # :97
    t2 = linear(hid, p[prefix + "linear2.weight"], p[prefix + "linear2.bias"])
    return tgt + _drop(t2, mk.get("drop3"), p_drop)                                  # :99


def memory_attention(p: Dict[str, Tensor], curr: Tensor, memory: Tensor,
                     curr_pos: Optional[Tensor], memory_pos: Tensor,
                     num_obj_ptr_tokens: int = 0, num_layers: int = NUM_LAYERS,
                     masks: Optional[list] = None, p_drop: float = 0.0) -> Tensor:
    """``MemoryAttention.forward`` (memory_attention.py:119-169).

    curr, curr_pos: [N, B, 256]; memory, memory_pos: [M, B, 64]; returns [N, B, 256].
    ``p`` uses the reference's state_dict keys (``layers.{i}.self_attn.q_proj.weight`` ...).
    """
    if curr.shape[1] != memory.shape[1]:
        raise AssertionError("Batch size must be the same for curr and memory")  # :135-137
    out = curr
    if curr_pos is not None:
        out = out + 0.1 * curr_pos  # pos_enc_at_input (:140-141)
    out = out.transpose(0, 1)
    qpos = curr_pos.transpose(0, 1) if curr_pos is not None else None
    mem = memory.transpose(0, 1)
    mpos = memory_pos.transpose(0, 1)
    for i in range(num_layers):
        out = memory_attention_layer(p, f"layers.{i}.", out, mem, mpos, qpos, num_obj_ptr_tokens,
                                     masks=masks[i] if masks is not None else None, p_drop=p_drop)
    out = layer_norm(out, p["norm.weight"], p["norm.bias"])
    return out.transpose(0, 1)


def init_params(seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random-init parameters with nn.Linear / nn.LayerNorm default statistics, keyed like the
    reference state_dict (106 tensors, 5 922 304 parameters)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}

    def lin(name, out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        p[name + ".weight"] = ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound).to(dtype)
        p[name + ".bias"] = ((torch.rand(out_f, generator=g) * 2 - 1) * bound).to(dtype)

    for i in range(NUM_LAYERS):
        pre = f"layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(pre + "self_attn." + nm, D_MODEL, D_MODEL)
        lin(pre + "cross_attn_image.q_proj", D_MODEL, D_MODEL)
        lin(pre + "cross_attn_image.k_proj", D_MODEL, KV_IN_DIM)
        lin(pre + "cross_attn_image.v_proj", D_MODEL, KV_IN_DIM)
        lin(pre + "cross_attn_image.out_proj", D_MODEL, D_MODEL)
        lin(pre + "linear1", DIM_FF, D_MODEL)
        lin(pre + "linear2", D_MODEL, DIM_FF)
        for nm in ("norm1", "norm2", "norm3"):
            p[pre + nm + ".weight"] = torch.ones(D_MODEL, dtype=dtype)
            p[pre + nm + ".bias"] = torch.zeros(D_MODEL, dtype=dtype)
    p["norm.weight"] = torch.ones(D_MODEL, dtype=dtype)
    p["norm.bias"] = torch.zeros(D_MODEL, dtype=dtype)
    return p


def reference_init_params(seed: int = 0) -> Dict[str, Tensor]:
    """The reference's OWN random initialisation, reproduced without the reference: the nn.Linear layers are created
    under ``torch.manual_seed(seed)`` in the order the reference constructors create them when the stack is built as
    ``oracle/ref_shim.build_memory_attention`` does (self-attention RoPEAttention first, then the cross-attention one --
    q, k, v, out each, transformer.py:213-216 -- then linear1, linear2 of MemoryAttentionLayer, memory_attention.py:39-41),
    and ``get_clones`` deep-copies that layer (sam2_utils.py:77-78): ALL FOUR LAYERS START IDENTICAL.  LayerNorms are
    (1, 0).  ``tests/test_oracle_golden.py`` checks this against the reference's state_dict when the reference is present."""
    torch.manual_seed(seed)
    lin = {}
    for pre, kv in (("self_attn.", D_MODEL), ("cross_attn_image.", KV_IN_DIM)):
        lin[pre + "q_proj"] = torch.nn.Linear(D_MODEL, D_MODEL)
        lin[pre + "k_proj"] = torch.nn.Linear(kv, D_MODEL)
        lin[pre + "v_proj"] = torch.nn.Linear(kv, D_MODEL)
        lin[pre + "out_proj"] = torch.nn.Linear(D_MODEL, D_MODEL)
    lin["linear1"] = torch.nn.Linear(D_MODEL, DIM_FF)
    lin["linear2"] = torch.nn.Linear(DIM_FF, D_MODEL)
    p: Dict[str, Tensor] = {}
    for i in range(NUM_LAYERS):
        pre = f"layers.{i}."
        for nm in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj",
                   "cross_attn_image.q_proj", "cross_attn_image.k_proj", "cross_attn_image.v_proj",
                   "cross_attn_image.out_proj", "linear1", "linear2"):
            p[pre + nm + ".weight"] = lin[nm].weight.detach().clone()
            p[pre + nm + ".bias"] = lin[nm].bias.detach().clone()
        for nm in ("norm1", "norm2", "norm3"):
            p[pre + nm + ".weight"] = torch.ones(D_MODEL)
            p[pre + nm + ".bias"] = torch.zeros(D_MODEL)
    p["norm.weight"] = torch.ones(D_MODEL)
    p["norm.bias"] = torch.zeros(D_MODEL)
    return p


def random_inputs(grid: int, batch: int, n_frames: int, n_ptr: int, seed: int = 1234) -> Dict[str, Tensor]:
    """Well-conditioned synthetic inputs of SURVEY.md section 8d: N(0,1) features, 0.7 N(0,1) positional encodings,
    N(0,1) upstream gradient, from one seeded CPU generator (bit-reproducible for a given torch build)."""
    g = torch.Generator().manual_seed(seed)
    n, m = grid * grid, n_frames * grid * grid + n_ptr
    return dict(curr=torch.randn(n, batch, D_MODEL, generator=g), curr_pos=torch.randn(n, batch, D_MODEL, generator=g) * 0.7,
                memory=torch.randn(m, batch, KV_IN_DIM, generator=g), memory_pos=torch.randn(m, batch, KV_IN_DIM, generator=g) * 0.7,
                grad_out=torch.randn(n, batch, D_MODEL, generator=g))


def core_attention(q: Tensor, k: Tensor, v: Tensor, num_k_exclude_rope: int,
                   scale: Optional[float] = None) -> Tensor:
    """Just the kernel-level op: rotate (q, first M-P keys) then softmax(q k^T * scale) v.
    q: [B, N, 256]; k, v: [B, M, 256] (already projected)."""
    n, m = q.shape[1], k.shape[1]
    cos, sin = (t.to(q.device) for t in axial_rope_table(n, q.shape[-1], dtype=q.dtype))
    nk = m - num_k_exclude_rope
    qr = apply_axial_rope(q, cos, sin)
    kr = torch.cat([apply_axial_rope(k[:, :nk], cos, sin), k[:, nk:]], dim=1) if nk > 0 else k
    sc = (1.0 / math.sqrt(q.shape[-1])) if scale is None else scale
    a = torch.softmax((qr @ kr.transpose(1, 2)) * sc, dim=-1)
    return a @ v
