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
