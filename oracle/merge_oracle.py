"""CPU oracle for the logit-producer side of the mask loss (TEST INFRASTRUCTURE ONLY) -- SURVEY.md section 8f rank 2.

Between the mask decoder and the loss the reference does, per frame:

1. ``F.interpolate(low_res_multimasks.float(), size=(image_size, image_size), mode="bilinear",
   align_corners=False)`` -- sam2_video/model/modeling/sam2_base.py:393-399 (4x up-sampling of the per-object logits);
2. ``merge_object_results_to_category`` -- sam2_video/utils/masks.py:53-212, called at
   sam2_video/model/sam2model.py:173-177: mask logits of the objects of one category are merged by a pixel-wise max
   (``_grouped_max`` :102-116, empty category -> zeros), IoU predictions by an average weighted with the objects'
   probability mass ``sum sigmoid(high-res logits)`` (``_area_weights_from_masks`` :92-100 -- NOT detached, so the IoU
   loss back-propagates into the logits through the weights -- and ``_grouped_weighted_avg`` :118-145);
3. ``MultiStepMultiMasksAndIous.forward`` on the merged per-category tensors (oracle/losses_oracle.py).

This file restates 1 and 2 with explicit index arithmetic (no ``F.interpolate``) and chains them into
``losses_oracle.multistep_loss``; everything is differentiable through torch autograd.

Parity status: PINNED against the unmodified reference functions executed in the build container
(``oracle/make_golden.py::golden_merged`` -> ``tests/golden/merged_*.npz``; ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import losses_oracle as lo

Tensor = torch.Tensor


def upsample_bilinear_x4(low: Tensor) -> Tensor:
    """``F.interpolate(..., mode="bilinear", align_corners=False)`` from [.., s, s] to [.., 4s, 4s]
    (sam2_base.py:393-399).  Destination pixel p samples the source at ``(p + 0.5) / 4 - 0.5`` clamped below at 0;
    i0 = floor, i1 = min(i0 + 1, s - 1), weight of i1 = fractional part (ATen ``area_pixel_compute_source_index``)."""
    s_h, s_w = low.shape[-2], low.shape[-1]

    def taps(s):
        p = torch.arange(4 * s, dtype=torch.float64)
        src = ((p + 0.5) / 4 - 0.5).clamp(min=0)
        i0 = src.floor().long()
        i1 = (i0 + 1).clamp(max=s - 1)
        lam = (src - i0).to(low.dtype)
        return i0, i1, lam

    y0, y1, ly = taps(s_h)
    x0, x1, lx = taps(s_w)
    rows = low[..., y0, :] * (1 - ly)[:, None] + low[..., y1, :] * ly[:, None]
    return rows[..., :, x0] * (1 - lx) + rows[..., :, x1] * lx


def category_groups(obj_to_cat: Sequence[int], num_categories: int) -> List[List[int]]:
    """masks.py:86-90: objects of each category in increasing object index."""
    groups: List[List[int]] = [[] for _ in range(num_categories)]
    for i, c in enumerate(obj_to_cat):
        groups[int(c)].append(int(i))
    return groups


def merge_frame(hi: Tensor, pred_ious: Tensor, groups: List[List[int]]):
    """One frame of ``merge_object_results_to_category`` (masks.py:147-212) for the two keys the loss reads.
    hi: [Nobj, 1, S, S] high-res logits; pred_ious: [Nobj, K].  Returns ([C, 1, S, S], [C, K])."""
    w = torch.sigmoid(hi).sum(dim=(1, 2, 3))                         # :92-100
    xs, ious = [], []
    for idxs in groups:
        if len(idxs) == 0:                                          # :111-112, :134-136
            xs.append(hi.new_zeros(hi.shape[1:]))
            ious.append(pred_ious.new_zeros(pred_ious.shape[1:]))
            continue
        xs.append(hi[idxs].max(dim=0).values)                       # :114 (gradient to the first maximal index)
        sw = w[idxs].view(-1, *([1] * (pred_ious.dim() - 1)))
        den = sw.sum(dim=0)
        if bool(torch.all(den == 0)):                               # :140-141
            ious.append(pred_ious[idxs].mean(dim=0))
        else:
            ious.append((pred_ious[idxs] * sw).sum(dim=0) / den)    # :143
    return torch.stack(xs, dim=0), torch.stack(ious, dim=0)


def merged_multistep_loss(low_res: Sequence[Tensor], pred_ious: Sequence[Tensor], obj_to_cat: Sequence[int],
                          num_categories: int, targets: Tensor, weight_dict: Dict[str, float], **loss_kw) -> Dict[str, Tensor]:
    """low_res[f]: [Nobj, 1, s, s]; pred_ious[f]: [Nobj, 1]; targets: [T, C, 4s, 4s] -> the loss dict of
    ``MultiStepMultiMasksAndIous`` on the merged per-category predictions."""
    groups = category_groups(obj_to_cat, num_categories)
    xs, ious = [], []
    for f in range(len(low_res)):
        x, iou = merge_frame(upsample_bilinear_x4(low_res[f].float()), pred_ious[f], groups)
        xs.append(x)
        ious.append(iou)
    return lo.multistep_loss(xs, targets, ious, weight_dict, **loss_kw)
