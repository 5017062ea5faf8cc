"""Deterministic analytic input generators shared by the golden-vector script and the tests
(TEST INFRASTRUCTURE ONLY).  ``det(shape, a, b, s)[i] = sin(i*a + b) * s`` over the flat index,
evaluated in fp64 and cast -- the recipe SURVEY.md section 8c used for its hand-checked values."""
from __future__ import annotations

import math
from typing import Dict

import torch


def det(shape, a: float, b: float, s: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    n = 1
    for d in shape:
        n *= d
    i = torch.arange(n, dtype=torch.float64)
    return (torch.sin(i * a + b) * s).reshape(shape).to(dtype)


def det_params(named_shapes, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """k-th tensor (k from 1, reference ``named_parameters()`` order): matrices
    det(shape, .013k+.1, k, 1/sqrt(fan_in)), LayerNorm weights 1 + det(shape, .7, k, .1),
    biases det(shape, .9, k, .05)."""
    out = {}
    for k, (name, shape) in enumerate(named_shapes, start=1):
        if len(shape) == 2:
            out[name] = det(shape, 0.013 * k + 0.1, k, 1.0 / math.sqrt(shape[1]), dtype)
        elif "norm" in name and name.endswith("weight"):
            out[name] = 1 + det(shape, 0.7, k, 0.1, dtype)
        else:
            out[name] = det(shape, 0.9, k, 0.05, dtype)
    return out


def param_shapes():
    """(name, shape) in the reference's ``named_parameters()`` order."""
    shapes = []
    for i in range(4):
        pre = f"layers.{i}."
        for att, kv in (("self_attn", 256), ("cross_attn_image", 64)):
            for nm, inf in (("q_proj", 256), ("k_proj", kv), ("v_proj", kv), ("out_proj", 256)):
                shapes.append((f"{pre}{att}.{nm}.weight", (256, inf)))
                shapes.append((f"{pre}{att}.{nm}.bias", (256,)))
        shapes.append((pre + "linear1.weight", (2048, 256)))
        shapes.append((pre + "linear1.bias", (2048,)))
        shapes.append((pre + "linear2.weight", (256, 2048)))
        shapes.append((pre + "linear2.bias", (256,)))
        for nm in ("norm1", "norm2", "norm3"):
            shapes.append((f"{pre}{nm}.weight", (256,)))
            shapes.append((f"{pre}{nm}.bias", (256,)))
    shapes.append(("norm.weight", (256,)))
    shapes.append(("norm.bias", (256,)))
    return shapes


def attention_inputs(grid: int, batch: int, n_frames: int, n_ptr_tokens: int, dtype=torch.float32):
    n = grid * grid
    m = n_frames * n + n_ptr_tokens
    return dict(
        curr=det((n, batch, 256), 0.21, 0.3, 1.0, dtype),
        curr_pos=det((n, batch, 256), 0.17, 1.3, 1.0, dtype),
        memory=det((m, batch, 64), 0.29, 2.3, 1.0, dtype),
        memory_pos=det((m, batch, 64), 0.31, 3.3, 1.0, dtype),
        grad_out=det((n, batch, 256), 0.41, 0.9, 1.0, dtype),
    )


def loss_inputs(t: int, c: int, s: int, clear=((1, 2),), dtype=torch.float32):
    logits = det((t, c, 1, s, s), 0.37, 0.11, 4.0, dtype)
    i = torch.arange(t * c * s * s, dtype=torch.int64)
    targets = (((i * 7) % 5) < 2).reshape(t, c, s, s)
    for (ft, fc) in clear:
        if ft < t and fc < c:
            targets[ft, fc] = False
    iou_pred = det((t, c, 1), 1.3, 0.5, 0.5, dtype) + 0.5
    return logits, targets, iou_pred


def multimask_loss_inputs(t: int, c: int, m: int, s: int, dtype=torch.float32):
    """M masks per channel (losses.py:143-238 with src_masks [C, M, H, W]): logits [t, c, m, s, s], IoU heads [t, c, m]."""
    _, targets, _ = loss_inputs(t, c, s)
    return det((t, c, m, s, s), 0.43, 0.19, 4.0, dtype), targets, det((t, c, m), 1.1, 0.3, 0.5, dtype) + 0.5


def merged_inputs(t: int, n_obj: int, c: int, s: int, clear=((1, 0),), dtype=torch.float32):
    """Producer-side fixture (oracle/merge_oracle.py): low-res logits [t, n_obj, 1, s, s], per-object IoU predictions
    [t, n_obj, 1], object -> category map with the LAST category left without objects, targets [t, c, 4s, 4s]."""
    low = det((t, n_obj, 1, s, s), 0.53, 0.27, 5.0, dtype)
    iou_pred = det((t, n_obj, 1), 1.7, 0.4, 0.45, dtype) + 0.5
    obj_to_cat = [(5 * i + 1) % max(c - 1, 1) for i in range(n_obj)]
    S = 4 * s
    i = torch.arange(t * c * S * S, dtype=torch.int64)
    targets = (((i * 11) % 7) < 2).reshape(t, c, S, S)
    for (ft, fc) in clear:
        if ft < t and fc < c:
            targets[ft, fc] = False
    return low, iou_pred, obj_to_cat, targets


def bank_scenarios():
    """(tag, kwargs) of the memory-bank assembly fixtures: frame index, conditioning / tracked frames present,
    train vs eval (stride, pointers in the past only), conditioning-frame limit, reverse tracking."""
    return [
        ("train_f5", dict(frame_idx=5, cond=[0], non_cond=[1, 2, 3, 4], num_frames=8, training=True)),
        ("train_f12_full", dict(frame_idx=12, cond=[0], non_cond=list(range(1, 12)), num_frames=20, training=True)),
        ("eval_stride2_f9", dict(frame_idx=9, cond=[0, 4], non_cond=[1, 2, 3, 5, 6, 7, 8], num_frames=16, training=False, stride=2)),
        ("eval_maxcond2_f10", dict(frame_idx=10, cond=[0, 3, 7, 14], non_cond=[8, 9], num_frames=16, training=False, max_cond=2)),
        ("reverse_f4", dict(frame_idx=4, cond=[9], non_cond=[5, 6, 7, 8], num_frames=10, training=False, reverse=True)),
    ]


def bank_inputs(cond, non_cond, b=2, h=6, w=6, c=256, md=64):
    """Deterministic per-frame outputs ({maskmem_features [B, md, H, W], maskmem_pos_enc [[B, md, H, W]], obj_ptr [B, C]})
    and the trainable tensors the assembly reads (maskmem_tpos_enc [7,1,1,md], obj_ptr_tpos_proj weight / bias)."""
    def frame(t):
        return {"maskmem_features": det((b, md, h, w), 0.11 + 0.01 * t, 0.3 * t, 1.0),
                "maskmem_pos_enc": [det((b, md, h, w), 0.07 + 0.02 * t, 1.1 * t, 0.7)],
                "obj_ptr": det((b, c), 0.19 + 0.03 * t, 0.5 * t, 1.0)}
    od = {"cond_frame_outputs": {t: frame(t) for t in cond}, "non_cond_frame_outputs": {t: frame(t) for t in non_cond}}
    return od, det((7, 1, 1, md), 0.37, 0.2, 0.05), det((md, c), 0.23, 0.4, 1.0 / math.sqrt(c)), det((md,), 0.9, 0.1, 0.05)
