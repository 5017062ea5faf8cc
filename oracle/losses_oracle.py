"""CPU oracle for the per-frame multi-object mask loss (TEST INFRASTRUCTURE ONLY).

Closed-form restatement of ``sam2_video/model/losses.py`` of the reference: per frame and per
valid channel six reductions (sum focal, sum p*t, sum p, sum t, |pred & gt|, |pred | gt|) and
the scalar algebra that follows.  It does not call ``binary_cross_entropy_with_logits`` or any
nn.Module; it is differentiable through autograd and also provides the analytic gradient
(:func:`multistep_loss_grad`) the CUDA backward kernel is derived from.

Only ``tests/``, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / reference legs may
import this module.  Parity status: PINNED against the unmodified reference executed through
``oracle/ref_shim.py`` (fixtures in tests/golden/, generator oracle/make_golden.py), including
the hand-checked values recorded in SURVEY.md section 8c.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

Tensor = torch.Tensor
CORE_LOSS_KEY = "total_loss"  # losses.py:17


def _softplus_neg_abs(x: Tensor) -> Tensor:
    return torch.log1p(torch.exp(-x.abs()))


def bce_with_logits(x: Tensor, t: Tensor, pos_weight: Optional[Tensor] = None) -> Tensor:
    """Element-wise BCE-with-logits (what losses.py:45 and :300-305 obtain from torch):
    ``(1-t)*x + (1 + (pw-1)*t) * softplus(-x)`` in its overflow-free form."""
    sp = torch.clamp(-x, min=0) + _softplus_neg_abs(x)  # softplus(-x)
    if pos_weight is None:
        return (1 - t) * x + sp
    return (1 - t) * x + (1 + (pos_weight - 1) * t) * sp


def frame_channel_sums(x: Tensor, t: Tensor, alpha: float, gamma: float) -> Dict[str, Tensor]:
    """x, t: [C, H*W] (x already divided by the temperature).  The six per-channel reductions.

    focal: losses.py:45-53; dice terms: :21-31; IoU areas: :63-66."""
    p = torch.sigmoid(x)
    ce = bce_with_logits(x, t)
    p_t = p * t + (1 - p) * (1 - t)
    fl = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        fl = (alpha * t + (1 - alpha) * (1 - t)) * fl
    pred = x > 0
    gt = t > 0
    return dict(
        focal=fl.sum(-1), pt=(p * t).sum(-1), p=p.sum(-1), t=t.sum(-1),
        inter=(pred & gt).sum(-1).to(x.dtype), union=(pred | gt).sum(-1).to(x.dtype),
    )


def multistep_loss(logits: Sequence[Tensor], targets: Tensor, pred_ious: Sequence[Tensor],
                   weight_dict: Dict[str, float], focal_alpha: float = 0.25,
                   focal_gamma: float = 2.0, iou_use_l1_loss: bool = False,
                   logit_temperature: float = 1.0) -> Dict[str, Tensor]:
    """``MultiStepMultiMasksAndIous.forward`` (losses.py:112-248) for the shape the training
    wrapper produces: one step per frame, one mask per channel (sam2model.py:472-476),
    ``pred_obj_scores=False`` so loss_class == 0 (:186-193).

    logits[f]: [C, 1, H, W]; targets: [T, C, H, W] (bool / 0-1); pred_ious[f]: [C, 1].
    Per frame: valid = channels with any foreground (:149-151), ``ValueError`` if none (:153-161),
    Nv = #valid (:174); focal_c = mean over pixels / Nv (:55-56); dice_c = (1 - (2 sum pt + 1) /
    (sum p + sum t + 1)) / Nv (:31-33); iou_c = |pred - I/max(U,1)| or squared, / Nv (:67-75);
    losses summed over channels (:231-233) and over frames (:116-119); total = sum_k w_k loss_k
    over non-zero weights (:240-248).
    """
    if len(logits) != len(targets):
        raise AssertionError("len(outs_batch) != len(targets_batch)")  # :113
    dt = logits[0].dtype
    out = {k: torch.zeros((), dtype=dt) for k in ("loss_mask", "loss_dice", "loss_iou", "loss_class")}
    for f in range(len(logits)):
        x = logits[f]
        c = x.shape[0]
        hw = x.shape[-1] * x.shape[-2]
        t = targets[f].reshape(c, hw).to(dt)
        valid = t.sum(-1) > 0
        if not bool(valid.any()):
            raise ValueError("No valid masks")
        nv = float(valid.sum())
        xs = x.reshape(c, hw)[valid] / logit_temperature
        s = frame_channel_sums(xs, t[valid], focal_alpha, focal_gamma)
        focal = s["focal"] / hw / nv
        dice = (1 - (2 * s["pt"] + 1) / (s["p"] + s["t"] + 1)) / nv
        actual = s["inter"] / torch.clamp(s["union"], min=1.0)
        diff = pred_ious[f].reshape(c)[valid] - actual.detach()
        iou = (diff.abs() if iou_use_l1_loss else diff ** 2) / nv
        out["loss_mask"] = out["loss_mask"] + focal.sum()
        out["loss_dice"] = out["loss_dice"] + dice.sum()
        out["loss_iou"] = out["loss_iou"] + iou.sum()
    total = torch.zeros((), dtype=dt)
    for k, w in weight_dict.items():
        if k not in out:
            raise ValueError(f"loss doesn't compute {k}")
        if w != 0:
            total = total + out[k] * w
    out[CORE_LOSS_KEY] = total
    return out


def multistep_loss_grad(logits: Tensor, targets: Tensor, pred_ious: Tensor,
                        weight_dict: Dict[str, float], focal_alpha: float = 0.25,
                        focal_gamma: float = 2.0, iou_use_l1_loss: bool = False,
                        logit_temperature: float = 1.0):
    """Analytic d total / d logits and d total / d pred_ious (SURVEY.md section 8a closed form).

    logits: [T, C, H*W]; targets same shape; pred_ious: [T, C].  Returns (dlogits, dious).
    d focal/dx = alpha_t [ (p - t) q^g + ce * g * q^(g-1) * (1 - 2t) p (1 - p) ],  q = 1 - p_t;
    d dice/dx  = -[ 2 t (D + 1) - (Nn + 1) ] / (D + 1)^2 * p (1 - p), Nn = 2 sum pt, D = sum p + sum t.
    """
    tt, c, hw = logits.shape
    dt = logits.dtype
    t = targets.to(dt)
    x = logits / logit_temperature
    valid = t.sum(-1) > 0  # [T, C]
    nv = valid.sum(-1, keepdim=True).to(dt)  # [T, 1]
    p = torch.sigmoid(x)
    ce = bce_with_logits(x, t)
    q = 1 - (p * t + (1 - p) * (1 - t))
    a_t = (focal_alpha * t + (1 - focal_alpha) * (1 - t)) if focal_alpha >= 0 else torch.ones_like(t)
    g = focal_gamma
    dfocal = a_t * ((p - t) * q ** g + ce * g * q ** (g - 1) * (1 - 2 * t) * p * (1 - p))
    nn_ = 2 * (p * t).sum(-1, keepdim=True)
    dd = p.sum(-1, keepdim=True) + t.sum(-1, keepdim=True)
    ddice = -(2 * t * (dd + 1) - (nn_ + 1)) / (dd + 1) ** 2 * p * (1 - p)
    w_m, w_d, w_i = weight_dict["loss_mask"], weight_dict["loss_dice"], weight_dict["loss_iou"]
    dx = (w_m * dfocal / hw + w_d * ddice) / (nv.unsqueeze(-1) * logit_temperature)
    dx = dx * valid.unsqueeze(-1)
    inter = ((x > 0) & (t > 0)).sum(-1).to(dt)
    union = ((x > 0) | (t > 0)).sum(-1).to(dt)
    diff = pred_ious - inter / torch.clamp(union, min=1.0)
    di = (torch.sign(diff) if iou_use_l1_loss else 2 * diff) * w_i / nv
    di = di * valid
    return dx, di


def bce_category_loss(logits: Sequence[Tensor], targets: Tensor,
                      pos_weight: Optional[Tensor] = None, reduction: str = "mean",
                      logit_temperature: float = 1.0) -> Dict[str, Tensor]:
    """``BCECategoryLoss.forward`` (losses.py:308-372): per frame keep channels with foreground
    (:342-344), scale by 1/T (:346), BCE-with-logits with optional per-channel pos_weight
    (:348-364) reduced by mean / sum over the kept elements, then the mean over frames (:368).
    An all-empty frame gives the mean of an empty tensor = NaN, exactly like the reference."""
    if len(logits) != len(targets):
        raise AssertionError("Mismatched sequence lengths")
    dt = logits[0].dtype
    total = torch.zeros((), dtype=dt)
    for f in range(len(logits)):
        x = logits[f]
        if x.dim() == 4 and x.shape[1] == 1:
            x = x[:, 0]
        elif x.dim() != 3:
            raise ValueError("Unexpected logits shape for BCECategoryLoss")
        t = targets[f]
        if t.dim() != 3:
            raise ValueError("Unexpected target shape for BCECategoryLoss")
        valid = t.reshape(t.shape[0], -1).sum(-1) > 0
        xs = x[valid] / logit_temperature
        ts = t[valid].to(dt)
        pw = None
        if pos_weight is not None:
            pw = pos_weight.to(dt).reshape(-1, 1, 1)
            if pw.shape[0] != xs.shape[0]:  # the reference compares AFTER filtering logits (:359-362)
                raise ValueError("pos_weight length does not match number of classes")
            pw = pw[valid]
        el = bce_with_logits(xs, ts, pw)
        if reduction == "mean":
            fl = el.mean()
        elif reduction == "sum":
            fl = el.sum()
        else:
            raise ValueError("reduction must be 'mean' or 'sum' for a scalar training loss")
        total = total + fl
    total = total / max(len(logits), 1)
    return {"loss_bce": total, CORE_LOSS_KEY: total}


# ---- stand-alone functional forms (losses.py:20-76), restated on top of the six per-channel sums -------------
def _sums(inputs: Tensor, targets: Tensor, lead: int, alpha: float = -1.0, gamma: float = 0.0):
    shape = tuple(inputs.shape[:lead])
    x = inputs.reshape(int(torch.Size(shape).numel()), -1)
    t = targets.reshape(x.shape[0], -1).to(x.dtype)
    s = frame_channel_sums(x, t, alpha, gamma)
    return {k: v.view(*shape) for k, v in s.items()}, x.shape[1]


def dice_loss(inputs: Tensor, targets: Tensor, num_objects: float, loss_on_multimask: bool = False) -> Tensor:
    """losses.py:20-34: 1 - (2 sum p t + 1) / (sum p + sum t + 1) per mask (multimask) or per row, / num_objects."""
    s, _ = _sums(inputs, targets, 2 if loss_on_multimask else 1)
    loss = 1 - (2 * s["pt"] + 1) / (s["p"] + s["t"] + 1)
    return loss / num_objects if loss_on_multimask else loss.sum() / num_objects


def sigmoid_focal_loss(inputs: Tensor, targets: Tensor, num_objects: float, alpha: float = 0.25, gamma: float = 2,
                       loss_on_multimask: bool = False) -> Tensor:
    """losses.py:37-57: multimask -> mean over H*W per mask / num_objects; else mean over dim 1, sum of the rest."""
    if loss_on_multimask:
        s, hw = _sums(inputs, targets, 2, alpha, gamma)
        return s["focal"] / hw / num_objects
    s, _ = _sums(inputs, targets, 1, alpha, gamma)
    return s["focal"].sum() / inputs.shape[1] / num_objects


def iou_loss(inputs: Tensor, targets: Tensor, pred_ious: Tensor, num_objects: float, loss_on_multimask: bool = False,
             use_l1_loss: bool = False) -> Tensor:
    """losses.py:60-76: |pred - I/max(U,1)| or its square, gradient to pred_ious only."""
    s, _ = _sums(inputs.detach(), (targets > 0), 2)
    actual = s["inter"] / torch.clamp(s["union"], min=1.0)
    d = pred_ious - actual
    loss = d.abs() if use_l1_loss else d * d
    return loss / num_objects if loss_on_multimask else loss.sum() / num_objects
