"""Memory-bank assembly for ``MemoryAttention`` on B200 -- the caller side of the hot path.

Replaces step 1 + the concatenation of ``SAM2Base._prepare_memory_conditioned_features``
(sam2_video/model/modeling/sam2_base.py:524-692) and its helpers ``select_closest_cond_frames`` / ``get_1d_sine_pe``
(sam2_video/model/modeling/sam2_utils.py:19-74).  Which past frames and object pointers enter the bank is decided on
the host with the reference's rules (same names, same arguments); the data movement -- per frame two
``flatten(2).permute(2, 0, 1)`` copies, the temporal-position add, the pointer split, two ``torch.cat`` -- is ONE launch
of ``sam2b200_bank_gather`` (csrc/bank.cu) that writes ``memory`` / ``memory_pos`` ``[M, B, 64]`` directly.

``memory_pos`` stays differentiable towards ``maskmem_tpos_enc`` and the object-pointer position projection
(``obj_ptr_tpos_proj``), the two trainable tensors it depends on (sam2_base.py:138-141, 608-610, 654-663); the memory
features and pointers also receive their gradient if they ask for it.  CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}


@dataclass
class BankConfig:
    """The ``SAM2Base`` attributes the assembly reads (sam2_base.py:28-100; values of configs/sam2/sam2.1_hiera_t.yaml)."""
    num_maskmem: int = 7
    hidden_dim: int = 256
    mem_dim: int = 64
    max_cond_frames_in_attn: int = -1
    memory_temporal_stride_for_eval: int = 1
    use_obj_ptrs_in_encoder: bool = True
    max_obj_ptrs_in_encoder: int = 16
    add_tpos_enc_to_obj_ptrs: bool = True
    proj_tpos_enc_in_obj_ptrs: bool = True
    use_signed_tpos_enc_to_obj_ptrs: bool = True
    only_obj_ptrs_in_the_past_for_eval: bool = True

    @classmethod
    def from_sam2_base(cls, m) -> "BankConfig":
        return cls(**{f: getattr(m, f) for f in cls.__dataclass_fields__ if hasattr(m, f)})


def select_closest_cond_frames(frame_idx: int, cond_frame_outputs: Dict[int, dict], max_cond_frame_num: int):
    """sam2_utils.py:19-61: up to ``max_cond_frame_num`` conditioning frames closest in time to ``frame_idx`` --
    the closest one before, the closest one at or after, then by distance.  Returns (selected, unselected)."""
    if max_cond_frame_num == -1 or len(cond_frame_outputs) <= max_cond_frame_num:
        return cond_frame_outputs, {}
    assert max_cond_frame_num >= 2, "we should allow using 2+ conditioning frames"
    selected: Dict[int, dict] = {}
    idx_before = max((t for t in cond_frame_outputs if t < frame_idx), default=None)
    if idx_before is not None:
        selected[idx_before] = cond_frame_outputs[idx_before]
    idx_after = min((t for t in cond_frame_outputs if t >= frame_idx), default=None)
    if idx_after is not None:
        selected[idx_after] = cond_frame_outputs[idx_after]
    num_remain = max_cond_frame_num - len(selected)
    for t in sorted((t for t in cond_frame_outputs if t not in selected), key=lambda x: abs(x - frame_idx))[:num_remain]:
        selected[t] = cond_frame_outputs[t]
    return selected, {t: v for t, v in cond_frame_outputs.items() if t not in selected}


def get_1d_sine_pe(pos_inds: Tensor, dim: int, temperature: float = 10000.0) -> Tensor:
    """sam2_utils.py:64-74 (a handful of scalars per frame: plain torch)."""
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32, device=pos_inds.device)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    pos_embed = pos_inds.unsqueeze(-1) / dim_t
    return torch.cat([pos_embed.sin(), pos_embed.cos()], dim=-1)


def select_bank_entries(cfg: BankConfig, frame_idx: int, output_dict: dict, num_frames: int, training: bool,
                        track_in_reverse: bool = False):
    """Host-side selection (sam2_base.py:551-596, 612-647): ([(t_pos, frame output)], [(temporal distance, obj_ptr)])."""
    selected, unselected = select_closest_cond_frames(frame_idx, output_dict["cond_frame_outputs"], cfg.max_cond_frames_in_attn)
    frames: List[Tuple[int, dict]] = [(0, out) for out in selected.values()]
    stride = 1 if training else cfg.memory_temporal_stride_for_eval
    for t_pos in range(1, cfg.num_maskmem):
        t_rel = cfg.num_maskmem - t_pos
        if t_rel == 1:
            prev_idx = frame_idx - t_rel if not track_in_reverse else frame_idx + t_rel
        elif not track_in_reverse:
            prev_idx = ((frame_idx - 2) // stride) * stride - (t_rel - 2) * stride
        else:
            prev_idx = -(-(frame_idx + 2) // stride) * stride + (t_rel - 2) * stride
        out = output_dict["non_cond_frame_outputs"].get(prev_idx, None)
        if out is None:
            out = unselected.get(prev_idx, None)
        if out is not None:
            frames.append((t_pos, out))
    pointers: List[Tuple[int, Tensor]] = []
    if cfg.use_obj_ptrs_in_encoder:
        sign = -1 if track_in_reverse else 1
        max_ptrs = min(num_frames, cfg.max_obj_ptrs_in_encoder)
        if not training and cfg.only_obj_ptrs_in_the_past_for_eval:
            ptr_cond = {t: o for t, o in selected.items() if (t >= frame_idx if track_in_reverse else t <= frame_idx)}
        else:
            ptr_cond = selected
        pointers = [(((frame_idx - t) * sign if cfg.use_signed_tpos_enc_to_obj_ptrs else abs(frame_idx - t)), o["obj_ptr"])
                    for t, o in ptr_cond.items()]
        for t_diff in range(1, max_ptrs):
            t = frame_idx + t_diff if track_in_reverse else frame_idx - t_diff
            if t < 0 or (num_frames is not None and t >= num_frames):
                break
            o = output_dict["non_cond_frame_outputs"].get(t, unselected.get(t, None))
            if o is not None:
                pointers.append((t_diff, o["obj_ptr"]))
    return frames, pointers


# Small host-built tensors that repeat from frame to frame (which tpos rows, the sine encoding of the pointer distances)
# are cached on the device: building them with torch.tensor(..., device=cuda) is a synchronous copy per call.
_IDX_CACHE: Dict[tuple, Tensor] = {}
_SINE_CACHE: Dict[tuple, Tensor] = {}


def _tpos_index(dev, rows: tuple) -> Tensor:
    key = (dev, rows)
    t = _IDX_CACHE.get(key)
    if t is None:
        if len(_IDX_CACHE) > 4096:
            _IDX_CACHE.clear()
        t = _IDX_CACHE[key] = torch.tensor(rows, device=dev)
    return t


def _pointer_sine_pe(dev, pos_list: tuple, t_diff_max: int, dim: int) -> Tensor:
    key = (dev, pos_list, t_diff_max, dim)
    t = _SINE_CACHE.get(key)
    if t is None:
        if len(_SINE_CACHE) > 4096:
            _SINE_CACHE.clear()
        t = _SINE_CACHE[key] = get_1d_sine_pe(torch.tensor(pos_list, dtype=torch.float32) / t_diff_max, dim=dim).to(dev)
    return t


class _BankGatherFn(torch.autograd.Function):
    """(memory, memory_pos) = gather(tpos_rows [S, 64], obj_pos [P, 64] | None, feats..., pos..., ptrs...)."""

    @staticmethod
    def forward(ctx, n_slots: int, n_ptrs: int, hidden_dim: int, tpos_rows, obj_pos, *tensors):
        lib = _lib.load()
        feats, pos, ptrs = tensors[:n_slots], tensors[n_slots:2 * n_slots], tensors[2 * n_slots:]
        ref = feats[0] if n_slots else ptrs[0]
        dev = ref.device
        b = ref.shape[0]
        hw = feats[0][0, 0].numel() if n_slots else 0
        per = hidden_dim // 64
        m = n_slots * hw + n_ptrs * per
        memory = torch.empty((m, b, 64), dtype=torch.float32, device=dev)
        memory_pos = torch.empty((m, b, 64), dtype=torch.float32, device=dev)
        tp = tpos_rows.detach().float().contiguous() if n_slots else None
        op = obj_pos.detach().float().contiguous() if obj_pos is not None else None
        rc = lib.sam2b200_bank_gather(
            _lib.ptr_array([t.data_ptr() for t in feats]) if n_slots else None,
            _lib.ptr_array([t.data_ptr() for t in pos]) if n_slots else None,
            _lib.ptr_array([tp[s].data_ptr() for s in range(n_slots)]) if n_slots else None, n_slots,
            _DT[feats[0].dtype] if n_slots else 0,
            _lib.ptr_array([t.data_ptr() for t in ptrs]) if n_ptrs else None, n_ptrs, _DT[ptrs[0].dtype] if n_ptrs else 0,
            op.data_ptr() if op is not None else None, memory.data_ptr(), memory_pos.data_ptr(), b, hw, 64, hidden_dim,
            torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "sam2b200_bank_gather")
        ctx.dims = (n_slots, n_ptrs, hw, b, per, hidden_dim)
        ctx.shapes = [tuple(t.shape) for t in tensors]
        ctx.dtypes = [t.dtype for t in tensors]
        ctx.has_obj_pos = obj_pos is not None
        return memory, memory_pos

    @staticmethod
    def backward(ctx, d_mem, d_pos):
        # Rarely-taken gradients (the bank is detached in training, sam2model.py:345-358) are view arithmetic in torch;
        # the one that always flows is d maskmem_tpos_enc = sum over tokens and batch of d memory_pos per frame.
        n_slots, n_ptrs, hw, b, per, c = ctx.dims
        needs = ctx.needs_input_grad
        d_tpos = d_obj = None
        sp = n_slots * hw
        if needs[3] and n_slots:
            d_tpos = d_pos[:sp].reshape(n_slots, hw * b, 64).sum(1)
        if ctx.has_obj_pos and needs[4] and n_ptrs:
            d_obj = d_pos[sp:].reshape(n_ptrs, per * b, 64).sum(1)
        grads: List[Optional[Tensor]] = []
        for k in range(len(ctx.shapes)):
            g = None
            if needs[5 + k]:
                if k < n_slots:                         # feats[k]: [B, 64, HW...] <- memory rows of slot k
                    g = d_mem[k * hw:(k + 1) * hw].permute(1, 2, 0).reshape(ctx.shapes[k]).to(ctx.dtypes[k])
                elif k < 2 * n_slots:                   # pos[k - n_slots]
                    s = k - n_slots
                    g = d_pos[s * hw:(s + 1) * hw].permute(1, 2, 0).reshape(ctx.shapes[k]).to(ctx.dtypes[k])
                else:                                   # ptrs[i]: [B, C] <- its C / 64 tokens
                    i = k - 2 * n_slots
                    g = d_mem[sp + i * per: sp + (i + 1) * per].permute(1, 0, 2).reshape(b, c).to(ctx.dtypes[k])
            grads.append(g)
        return (None, None, None, d_tpos, d_obj, *grads)


def _prepare(cfg: BankConfig, frame_idx: int, output_dict: dict, num_frames: int, maskmem_tpos_enc: Tensor,
             obj_ptr_tpos_proj=None, training: bool = True, track_in_reverse: bool = False):
    """Selection + validation + the small differentiable tensors (tpos rows, pointer positions) shared by both back ends."""
    if cfg.mem_dim != 64:
        raise _lib.Sam2B200Error("the B200 bank kernels are built for mem_dim = 64 (SAM2 memory encoder out_dim)")
    frames, pointers = select_bank_entries(cfg, frame_idx, output_dict, num_frames, training, track_in_reverse)
    if not frames and not pointers:
        raise ValueError("empty memory bank: no past frame outputs for this frame")
    dev = maskmem_tpos_enc.device
    feats = [o["maskmem_features"].to(dev, non_blocking=True) for _, o in frames]
    pos = [o["maskmem_pos_enc"][-1].to(dev) for _, o in frames]
    for t in feats + pos:
        if not t.is_cuda:
            raise _lib.Sam2B200Error("memory-bank tensors must be CUDA tensors: the B200 path has no CPU fallback")
    if feats and (feats[0].dtype not in _DT or any(t.dtype != feats[0].dtype for t in feats + pos)):
        feats, pos = [t.float() for t in feats], [t.float() for t in pos]
    feats = [t.contiguous() for t in feats]
    pos = [t.contiguous() for t in pos]
    # the reference's torch.cat / torch.stack raise on a frame with another batch size or resolution (an object added
    # mid-video, mixed resolutions); the gather kernel takes raw pointers, so the same check happens here
    if feats:
        shape0 = tuple(feats[0].shape)
        if len(shape0) < 3 or shape0[1] != cfg.mem_dim:
            raise ValueError(f"maskmem_features must be [B, {cfg.mem_dim}, H, W], got {shape0}")
        for (t_pos, _), f_, p_ in zip(frames, feats, pos):
            if tuple(f_.shape) != shape0 or tuple(p_.shape) != shape0:
                raise ValueError(f"memory frame at t_pos={t_pos}: maskmem_features {tuple(f_.shape)} / maskmem_pos_enc "
                                 f"{tuple(p_.shape)} do not match the first frame's {shape0}")
    tpos_rows = None
    if frames:
        idx = _tpos_index(dev, tuple(cfg.num_maskmem - t_pos - 1 for t_pos, _ in frames))            # sam2_base.py:608-610
        tpos_rows = maskmem_tpos_enc.reshape(cfg.num_maskmem, cfg.mem_dim).index_select(0, idx)
    ptrs: List[Tensor] = []
    obj_pos = None
    if pointers:
        pos_list, ptrs = zip(*pointers)
        ptrs = [p.to(dev) for p in ptrs]
        if ptrs[0].dtype not in _DT or any(p.dtype != ptrs[0].dtype for p in ptrs):
            ptrs = [p.float() for p in ptrs]
        ptrs = [p.contiguous() for p in ptrs]
        want = (feats[0].shape[0] if feats else ptrs[0].shape[0], cfg.hidden_dim)
        for t_diff, p_ in zip(pos_list, ptrs):
            if tuple(p_.shape) != want:
                raise ValueError(f"object pointer at t_diff={t_diff}: shape {tuple(p_.shape)}, expected {want}")
        if cfg.add_tpos_enc_to_obj_ptrs:                                                              # :654-663
            t_diff_max = min(num_frames, cfg.max_obj_ptrs_in_encoder) - 1
            tpos_dim = cfg.hidden_dim if cfg.proj_tpos_enc_in_obj_ptrs else cfg.mem_dim
            obj_pos = _pointer_sine_pe(dev, tuple(int(x) for x in pos_list), t_diff_max, tpos_dim)
            if obj_ptr_tpos_proj is not None:
                obj_pos = obj_ptr_tpos_proj(obj_pos)
            obj_pos = obj_pos.reshape(len(pos_list), cfg.mem_dim)
    return frames, ptrs, tpos_rows, obj_pos, feats, pos


def assemble_memory(cfg: BankConfig, frame_idx: int, output_dict: dict, num_frames: int, maskmem_tpos_enc: Tensor,
                    obj_ptr_tpos_proj=None, training: bool = True, track_in_reverse: bool = False):
    """``(memory [M, B, 64], memory_pos [M, B, 64], num_obj_ptr_tokens)`` for a frame that is not an initial
    conditioning frame -- the arguments ``SAM2Base.memory_attention`` is called with (sam2_base.py:695-709).

    ``output_dict`` is the reference's ``{"cond_frame_outputs": {t: out}, "non_cond_frame_outputs": {t: out}}`` with
    ``out["maskmem_features"] [B, 64, H, W]``, ``out["maskmem_pos_enc"][-1] [B, 64, H, W]``, ``out["obj_ptr"] [B, C]``;
    ``maskmem_tpos_enc [num_maskmem, 1, 1, 64]``; ``obj_ptr_tpos_proj``: the ``nn.Linear(C, 64)`` (or Identity)."""
    frames, ptrs, tpos_rows, obj_pos, feats, pos = _prepare(cfg, frame_idx, output_dict, num_frames, maskmem_tpos_enc,
                                                           obj_ptr_tpos_proj, training, track_in_reverse)
    memory, memory_pos = _BankGatherFn.apply(len(frames), len(ptrs), cfg.hidden_dim, tpos_rows, obj_pos, *feats, *pos, *ptrs)
    return memory, memory_pos, len(ptrs) * (cfg.hidden_dim // cfg.mem_dim)


@dataclass
class PackedBank:
    """The memory bank in the layout the fused MemoryAttention stack reads: batch-first, bf16,
    ``memk = memory + memory_pos`` (key source) and ``memv = memory`` (value source), both ``[B, M, 64]``.  Pass it as the
    ``memory`` argument of ``MemoryAttention`` (``memory_pos`` is then ignored): no fp32 ``[M, B, 64]`` tensors are
    materialised and the stack does not re-pack.

    The frame features and pointers are constants here (the training wrapper detaches them, sam2model.py:345-358); what stays
    differentiable is what ``memory_pos`` depends on: ``tpos_rows [n_slots, 64]`` (rows of maskmem_tpos_enc) and ``obj_pos
    [n_ptrs, 64]`` (projected pointer positions).  The stack's backward reduces its fp32 key-source gradient over tokens and
    objects itself and hands the two small gradients to autograd (nothing is rounded to bf16 on the way)."""
    memk: Tensor
    memv: Tensor
    num_obj_ptr_tokens: int
    tpos_rows: Optional[Tensor] = None
    obj_pos: Optional[Tensor] = None
    n_slots: int = 0
    hw: int = 0

    @property
    def shape(self):            # (M, B, 64), like the seq-first tensor it stands for (memory_attention.py:135-137 asserts on it)
        return (self.memk.shape[1], self.memk.shape[0], self.memk.shape[2])


def _pack(n_slots: int, n_ptrs: int, hidden_dim: int, tpos_rows, obj_pos, feats, pos, ptrs):
    """sam2b200_bank_gather_packed: (memk, memv) bf16 [B, M, 64]; no autograd (see PackedBank)."""
    lib = _lib.load()
    ref = feats[0] if n_slots else ptrs[0]
    dev = ref.device
    b = ref.shape[0]
    hw = feats[0][0, 0].numel() if n_slots else 0
    per = hidden_dim // 64
    m = n_slots * hw + n_ptrs * per
    memk = torch.empty((b, m, 64), dtype=torch.bfloat16, device=dev)
    memv = torch.empty((b, m, 64), dtype=torch.bfloat16, device=dev)
    tp = tpos_rows.detach().float().contiguous() if n_slots else None
    op = obj_pos.detach().float().contiguous() if obj_pos is not None else None
    rc = lib.sam2b200_bank_gather_packed(
        _lib.ptr_array([t.data_ptr() for t in feats]) if n_slots else None,
        _lib.ptr_array([t.data_ptr() for t in pos]) if n_slots else None,
        _lib.ptr_array([tp[s].data_ptr() for s in range(n_slots)]) if n_slots else None, n_slots,
        _DT[feats[0].dtype] if n_slots else 0,
        _lib.ptr_array([t.data_ptr() for t in ptrs]) if n_ptrs else None, n_ptrs, _DT[ptrs[0].dtype] if n_ptrs else 0,
        op.data_ptr() if op is not None else None, memk.data_ptr(), memv.data_ptr(), b, hw, 64, hidden_dim,
        torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "sam2b200_bank_gather_packed")
    return memk, memv, hw


def assemble_memory_packed(cfg: BankConfig, frame_idx: int, output_dict: dict, num_frames: int, maskmem_tpos_enc: Tensor,
                           obj_ptr_tpos_proj=None, training: bool = True, track_in_reverse: bool = False) -> PackedBank:
    """Same selection rules and inputs as :func:`assemble_memory`, but the bank is written ONCE, directly in the kernel
    layout of the fused stack (``PackedBank``): per frame the reference's ``[B, 64, H, W]`` memory-encoder outputs are
    transposed to token-major bf16 rows with the temporal position added on write, the pointer tokens appended behind
    them -- no fp32 ``[M, B, 64]`` pair, no ``memory + pos`` pass, no re-pack inside ``MemoryAttention``."""
    frames, ptrs, tpos_rows, obj_pos, feats, pos = _prepare(cfg, frame_idx, output_dict, num_frames, maskmem_tpos_enc,
                                                           obj_ptr_tpos_proj, training, track_in_reverse)
    with torch.no_grad():
        memk, memv, hw = _pack(len(frames), len(ptrs), cfg.hidden_dim, tpos_rows, obj_pos, feats, pos, ptrs)
    return PackedBank(memk, memv, len(ptrs) * (cfg.hidden_dim // cfg.mem_dim), tpos_rows=tpos_rows, obj_pos=obj_pos,
                      n_slots=len(frames), hw=hw)
