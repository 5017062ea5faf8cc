"""Mask loss fused with its producer side: low-res per-object logits in, loss dict out.

Host-side mirror of what the reference does between the mask decoder and the loss value, per frame:

* ``F.interpolate(low_res_multimasks.float(), size=(image_size, image_size), mode="bilinear", align_corners=False)``
  -- sam2_video/model/modeling/sam2_base.py:393-399;
* ``merge_object_results_to_category(previous_stages_out, obj_to_cat, num_categories)`` -- sam2_video/utils/masks.py:53-212,
  called at sam2_video/model/sam2model.py:173-177 (pixel-wise max of the logits over the objects of a category; IoU
  predictions averaged with the area weights ``sum(sigmoid(pred_masks_high_res))``, which are not detached);
* ``MultiStepMultiMasksAndIous.forward`` -- sam2_video/model/losses.py:112-248.

:class:`CategoryMergedMultiStepLoss` takes the *un-merged* per-frame dicts (low-res object logits + per-object IoU
predictions), ``obj_to_cat`` and ``num_categories`` and returns the same loss dict, without ever writing the
``[n_obj, 1, S, S]`` / ``[C, 1, S, S]`` high-resolution logits to memory (csrc/merge_loss.cu).  CUDA only: there is no
CPU fallback and a missing library raises.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
from torch import nn

from . import _lib
from . import ops as _ops
from .losses import CORE_LOSS_KEY, _raise_if_no_valid, _require_cuda, _stream_ptr, _targets_u8

_GROUP_CACHE: Dict[tuple, tuple] = {}


def category_groups(obj_to_cat: Sequence[int], num_categories: int, device) -> tuple:
    """CSR (offsets [C+1], members [n_obj]) of the objects of each category in increasing object index
    (masks.py:86-90), as int32 device tensors; cached per (mapping, device)."""
    key = (tuple(int(c) for c in obj_to_cat), int(num_categories), str(device))
    hit = _GROUP_CACHE.get(key)
    if hit is not None:
        return hit
    groups: List[List[int]] = [[] for _ in range(num_categories)]
    for i, c in enumerate(key[0]):
        if not 0 <= c < num_categories:
            raise IndexError(f"obj_to_cat[{i}] = {c} outside [0, {num_categories})")
        groups[c].append(i)
    if max((len(g) for g in groups), default=0) > 255:
        raise NotImplementedError("more than 255 objects in one category")
    offsets, members = [0], []
    for g in groups:
        members.extend(g)
        offsets.append(len(members))
    out = (torch.tensor(offsets, dtype=torch.int32, device=device), torch.tensor(members, dtype=torch.int32, device=device))
    if len(_GROUP_CACHE) > 256:
        _GROUP_CACHE.clear()
    _GROUP_CACHE[key] = out
    return out


class _MergedLossFn(torch.autograd.Function):
    """(losses[4], n_valid, cat_iou) = fused(obj_iou [T, n_obj], low_0 .. low_{T-1}); see sam2b200_merged_loss_fwd."""

    @staticmethod
    def forward(ctx, cfg, targets_u8, offsets, members, obj_iou, *low):
        lib = _lib.load()
        t = len(low)
        n_obj, s = low[0].shape[0], low[0].shape[-1]
        c = targets_u8.shape[1]
        dev = low[0].device
        ws = torch.empty(max(lib.sam2b200_merged_loss_workspace_bytes(t, c, n_obj, s), 4) // 4, dtype=torch.float32, device=dev)
        chan_sums = torch.empty(t, c, 6, dtype=torch.float32, device=dev)
        obj_area = torch.empty(t, n_obj, dtype=torch.float32, device=dev)
        cat_iou = torch.empty(t, c, dtype=torch.float32, device=dev)
        cat_w = torch.empty(t, c, dtype=torch.float32, device=dev)
        n_valid = torch.empty(t, dtype=torch.int32, device=dev)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        px = 16.0 * s * s
        with _ops._Timed("merged_loss_fwd", t * (c * px + 4.0 * n_obj * s * s)):
            rc = lib.sam2b200_merged_loss_fwd(
                _lib.ptr_array([x.data_ptr() for x in low]), targets_u8.data_ptr(), obj_iou.data_ptr(), offsets.data_ptr(),
                members.data_ptr(), ws.data_ptr(), chan_sums.data_ptr(), obj_area.data_ptr(), cat_iou.data_ptr(),
                cat_w.data_ptr(), n_valid.data_ptr(), losses.data_ptr(), t, c, n_obj, s, cfg["alpha"], cfg["gamma"],
                cfg["inv_temp"], int(cfg["iou_l1"]), _stream_ptr(dev))
        _lib.check(rc, "sam2b200_merged_loss_fwd")
        ctx.cfg = cfg
        ctx.dims = (t, c, n_obj, s)
        ctx.low_shapes = [tuple(x.shape) for x in low]
        ctx.save_for_backward(targets_u8, offsets, members, obj_iou, chan_sums, obj_area, cat_iou, cat_w, n_valid, *low)
        ctx.mark_non_differentiable(n_valid, cat_iou)
        return losses, n_valid, cat_iou

    @staticmethod
    def backward(ctx, g_losses, _g_nv, _g_iou):
        lib = _lib.load()
        targets_u8, offsets, members, obj_iou, chan_sums, obj_area, cat_iou, cat_w, n_valid, *low = ctx.saved_tensors
        t, c, n_obj, s = ctx.dims
        cfg = ctx.cfg
        dev = low[0].device
        g = g_losses.contiguous().float()
        dlow = torch.empty(t, n_obj, s, s, dtype=torch.float32, device=dev)
        d_iou = torch.empty(t, n_obj, dtype=torch.float32, device=dev)
        px = 16.0 * s * s
        with _ops._Timed("merged_loss_bwd", t * (c * px + 8.0 * n_obj * s * s)):
            rc = lib.sam2b200_merged_loss_bwd(
                _lib.ptr_array([x.data_ptr() for x in low]), _lib.ptr_array([dlow[f].data_ptr() for f in range(t)]),
                targets_u8.data_ptr(), obj_iou.data_ptr(), offsets.data_ptr(), members.data_ptr(), chan_sums.data_ptr(),
                obj_area.data_ptr(), cat_iou.data_ptr(), cat_w.data_ptr(), n_valid.data_ptr(), g.data_ptr(), d_iou.data_ptr(),
                t, c, n_obj, s, cfg["alpha"], cfg["gamma"], cfg["inv_temp"], int(cfg["iou_l1"]), _stream_ptr(dev))
        _lib.check(rc, "sam2b200_merged_loss_bwd")
        return (None, None, None, None, d_iou, *[dlow[f].view(ctx.low_shapes[f]) for f in range(t)])


class CategoryMergedMultiStepLoss(nn.Module):
    """``interpolate -> merge_object_results_to_category -> MultiStepMultiMasksAndIous`` in one fused op.

    ``forward(stages, obj_to_cat, num_categories, targets_batch)``: ``stages`` is the per-frame list the tracker produces
    *before* the merge, each dict holding ``"multistep_pred_multimasks"`` = ``[low-res logits [n_obj, 1, s, s]]`` and
    ``"multistep_pred_ious"`` = ``[[n_obj, 1]]`` (one step per frame, one mask per object -- what the training wrapper
    emits, sam2model.py:472-476); ``targets_batch``: ``[T, C, 4s, 4s]`` category masks.  Returns the reference's loss dict.
    Constructor arguments as ``MultiStepMultiMasksAndIous`` (losses.py:79-110); ``pred_obj_scores`` must stay off
    (the shipped configs, ``loss_class`` weight 0).
    """

    def __init__(self, weight_dict, focal_alpha=0.25, focal_gamma=2.0, supervise_all_iou=False, iou_use_l1_loss=False,
                 pred_obj_scores=False, focal_gamma_obj_score=0.0, focal_alpha_obj_score=-1, logit_temperature: float = 1.0,
                 check_valid: bool = True):
        super().__init__()
        self.weight_dict = weight_dict
        for k in ("loss_mask", "loss_dice", "loss_iou"):   # losses.py:96-98
            assert k in self.weight_dict
        if "loss_class" not in self.weight_dict:
            self.weight_dict["loss_class"] = 0.0
        if pred_obj_scores:
            raise NotImplementedError("pred_obj_scores=True (object-score focal term) is outside the fused producer path")
        if not (isinstance(logit_temperature, (int, float)) and logit_temperature > 0):
            raise ValueError("logit_temperature must be a positive float")   # losses.py:107-108
        self.focal_alpha, self.focal_gamma = focal_alpha, focal_gamma
        self.supervise_all_iou, self.iou_use_l1_loss = supervise_all_iou, iou_use_l1_loss
        self.logit_temperature = float(logit_temperature)
        self.check_valid = check_valid

    def forward(self, stages: List[Dict], obj_to_cat: Sequence[int], num_categories: int,
                targets_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        assert len(stages) == len(targets_batch)            # losses.py:113
        _require_cuda(targets_batch, "targets_batch")
        if targets_batch.shape[1] != num_categories:
            raise ValueError("targets_batch must hold one channel per category")
        low, ious = [], []
        for st in stages:
            a, b = st["multistep_pred_multimasks"], st["multistep_pred_ious"]
            assert len(a) == len(b)                          # losses.py:130-131
            if len(a) != 1:
                raise NotImplementedError("more than one correction step per frame")
            x = a[0]
            _require_cuda(x, "low-res mask logits")
            if x.dim() != 4 or x.shape[1] != 1 or x.shape[-1] != x.shape[-2]:
                raise NotImplementedError("expected square single-mask low-res logits [n_obj, 1, s, s]")
            if x.shape[0] != len(obj_to_cat):
                raise ValueError("obj_to_cat must name a category for every object")
            if 4 * x.shape[-1] != targets_batch.shape[-1] or targets_batch.shape[-1] != targets_batch.shape[-2]:
                raise ValueError("targets must be [T, C, 4s, 4s] for low-res logits [n_obj, 1, s, s]")
            x = x[:, 0]
            low.append(x if (x.dtype == torch.float32 and x.is_contiguous()) else x.contiguous().float())   # sam2_base.py:393
            ious.append(b[0].reshape(-1))
        offsets, members = category_groups(obj_to_cat, num_categories, targets_batch.device)
        cfg = dict(alpha=float(self.focal_alpha), gamma=float(self.focal_gamma), inv_temp=1.0 / self.logit_temperature,
                   iou_l1=bool(self.iou_use_l1_loss))
        obj_iou = torch.stack(ious).float().contiguous()
        losses4, n_valid, _ = _MergedLossFn.apply(cfg, _targets_u8(targets_batch), offsets, members, obj_iou, *low)
        if self.check_valid:
            _raise_if_no_valid(n_valid)
        losses = {"loss_mask": losses4[0], "loss_dice": losses4[1], "loss_iou": losses4[2], "loss_class": losses4[3]}
        total = 0.0
        for k, w in self.weight_dict.items():                # losses.py:240-248
            if k not in losses:
                raise ValueError(f"{type(self)} doesn't compute {k}")
            if w != 0:
                total = total + losses[k] * w
        losses[CORE_LOSS_KEY] = total
        return losses
