"""CPU oracle for the memory-bank assembly that feeds MemoryAttention (TEST INFRASTRUCTURE ONLY).

Restates, in plain torch ops, step 1 and the concatenation of
``SAM2Base._prepare_memory_conditioned_features`` (sam2_video/model/modeling/sam2_base.py:524-692) and its helpers
``select_closest_cond_frames`` / ``get_1d_sine_pe`` (sam2_video/model/modeling/sam2_utils.py:19-74): which past frames
enter the bank, their temporal position, the object-pointer tokens and their positional encoding, and the final
``memory`` / ``memory_pos`` tensors ``[M, B, 64]`` plus ``num_obj_ptr_tokens``.

Only ``tests/`` may import this module.  Parity status: PINNED against the unmodified reference method executed with
stub sub-modules through ``oracle/ref_shim.load_sam2_base()`` (fixtures tests/golden/bank_*.npz, generator
oracle/make_golden.py:golden_bank).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor


@dataclass
class BankConfig:
    """The SAM2Base attributes the assembly reads (defaults = configs/sam2/sam2.1_hiera_t.yaml + sam2_base.py:28-100)."""
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


def select_closest_cond_frames(frame_idx: int, cond_frame_outputs: Dict[int, dict], max_cond_frame_num: int):
    """sam2_utils.py:19-61: the closest conditioning frame before, the closest at-or-after, then by |distance|."""
    if max_cond_frame_num == -1 or len(cond_frame_outputs) <= max_cond_frame_num:
        return cond_frame_outputs, {}
    assert max_cond_frame_num >= 2
    selected: Dict[int, dict] = {}
    before = [t for t in cond_frame_outputs if t < frame_idx]
    if before:
        selected[max(before)] = cond_frame_outputs[max(before)]
    after = [t for t in cond_frame_outputs if t >= frame_idx]
    if after:
        selected[min(after)] = cond_frame_outputs[min(after)]
    rest = sorted((t for t in cond_frame_outputs if t not in selected), key=lambda x: abs(x - frame_idx))
    for t in rest[:max_cond_frame_num - len(selected)]:
        selected[t] = cond_frame_outputs[t]
    return selected, {t: v for t, v in cond_frame_outputs.items() if t not in selected}


def get_1d_sine_pe(pos_inds: Tensor, dim: int, temperature: float = 10000.0) -> Tensor:
    """sam2_utils.py:64-74."""
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    x = pos_inds.unsqueeze(-1) / dim_t
    return torch.cat([x.sin(), x.cos()], dim=-1)


def select_memory_frames(cfg: BankConfig, frame_idx: int, output_dict: dict, training: bool,
                         track_in_reverse: bool = False) -> Tuple[List[Tuple[int, Optional[dict]]], dict, dict]:
    """sam2_base.py:551-596: [(t_pos, frame output | None)] -- selected conditioning frames with t_pos 0, then the
    num_maskmem - 1 most recent frames (stride r in eval), oldest first."""
    selected, unselected = select_closest_cond_frames(frame_idx, output_dict["cond_frame_outputs"], cfg.max_cond_frames_in_attn)
    t_pos_and_prevs: List[Tuple[int, Optional[dict]]] = [(0, out) for out in selected.values()]
    stride = 1 if training else cfg.memory_temporal_stride_for_eval
    for t_pos in range(1, cfg.num_maskmem):
        t_rel = cfg.num_maskmem - t_pos
        if t_rel == 1:
            prev = frame_idx - t_rel if not track_in_reverse else frame_idx + t_rel
        elif not track_in_reverse:
            prev = ((frame_idx - 2) // stride) * stride - (t_rel - 2) * stride
        else:
            prev = -(-(frame_idx + 2) // stride) * stride + (t_rel - 2) * stride
        out = output_dict["non_cond_frame_outputs"].get(prev, None)
        if out is None:
            out = unselected.get(prev, None)
        t_pos_and_prevs.append((t_pos, out))
    return t_pos_and_prevs, selected, unselected


def select_object_pointers(cfg: BankConfig, frame_idx: int, output_dict: dict, num_frames: int, training: bool,
                           selected: dict, unselected: dict, track_in_reverse: bool = False) -> List[Tuple[int, Tensor]]:
    """sam2_base.py:612-647: [(temporal distance, obj_ptr [B, C])]."""
    sign = -1 if track_in_reverse else 1
    max_ptrs = min(num_frames, cfg.max_obj_ptrs_in_encoder)
    if not training and cfg.only_obj_ptrs_in_the_past_for_eval:
        ptr_cond = {t: o for t, o in selected.items() if (t >= frame_idx if track_in_reverse else t <= frame_idx)}
    else:
        ptr_cond = selected
    out = [(((frame_idx - t) * sign if cfg.use_signed_tpos_enc_to_obj_ptrs else abs(frame_idx - t)), o["obj_ptr"])
           for t, o in ptr_cond.items()]
    for t_diff in range(1, max_ptrs):
        t = frame_idx + t_diff if track_in_reverse else frame_idx - t_diff
        if t < 0 or (num_frames is not None and t >= num_frames):
            break
        o = output_dict["non_cond_frame_outputs"].get(t, unselected.get(t, None))
        if o is not None:
            out.append((t_diff, o["obj_ptr"]))
    return out


def assemble_memory(cfg: BankConfig, frame_idx: int, output_dict: dict, num_frames: int, maskmem_tpos_enc: Tensor,
                    obj_ptr_tpos_proj_weight: Optional[Tensor], obj_ptr_tpos_proj_bias: Optional[Tensor], training: bool,
                    track_in_reverse: bool = False) -> Tuple[Tensor, Tensor, int]:
    """(memory [M, B, mem_dim], memory_pos [M, B, mem_dim], num_obj_ptr_tokens) for a non-initial frame."""
    t_pos_and_prevs, selected, unselected = select_memory_frames(cfg, frame_idx, output_dict, training, track_in_reverse)
    mem, pos = [], []
    for t_pos, prev in t_pos_and_prevs:
        if prev is None:
            continue
        mem.append(prev["maskmem_features"].flatten(2).permute(2, 0, 1))                       # :602-603
        enc = prev["maskmem_pos_enc"][-1].flatten(2).permute(2, 0, 1)                          # :605-606
        pos.append(enc + maskmem_tpos_enc[cfg.num_maskmem - t_pos - 1])                        # :608-610
    n_ptr_tokens = 0
    if cfg.use_obj_ptrs_in_encoder:
        pos_and_ptrs = select_object_pointers(cfg, frame_idx, output_dict, num_frames, training, selected, unselected, track_in_reverse)
        if pos_and_ptrs:
            pos_list, ptrs = zip(*pos_and_ptrs)
            obj_ptrs = torch.stack(ptrs, dim=0)                                                # [P, B, C]
            b = obj_ptrs.shape[1]
            c, md = cfg.hidden_dim, cfg.mem_dim
            if cfg.add_tpos_enc_to_obj_ptrs:                                                   # :654-663
                t_max = min(num_frames, cfg.max_obj_ptrs_in_encoder) - 1
                obj_pos = get_1d_sine_pe(torch.tensor(pos_list, dtype=torch.float32) / t_max, dim=c if cfg.proj_tpos_enc_in_obj_ptrs else md)
                if cfg.proj_tpos_enc_in_obj_ptrs:
                    obj_pos = obj_pos @ obj_ptr_tpos_proj_weight.t() + obj_ptr_tpos_proj_bias
                obj_pos = obj_pos.unsqueeze(1).expand(-1, b, md)
            else:
                obj_pos = obj_ptrs.new_zeros(len(pos_list), b, md)
            if md < c:                                                                         # :666-672
                obj_ptrs = obj_ptrs.reshape(-1, b, c // md, md).permute(0, 2, 1, 3).flatten(0, 1)
                obj_pos = obj_pos.repeat_interleave(c // md, dim=0)
            mem.append(obj_ptrs)
            pos.append(obj_pos)
            n_ptr_tokens = obj_ptrs.shape[0]
    return torch.cat(mem, dim=0), torch.cat(pos, dim=0), n_ptr_tokens
