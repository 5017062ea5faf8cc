"""Drop-in replacements for ``sam2_video/model/losses.py`` of the reference, backed by the fused
sm_100a mask-loss kernels of libsam2b200.so (csrc/mask_loss.cu).

Same names, constructor arguments, ``forward(outs_batch, targets_batch) -> Dict[str, Tensor]``
contract, result keys and error behaviour as the reference (file:line cited per symbol); the
switch in ``SAM2LightningModule.__init__`` (sam2_video/training/trainer.py:67-94) is described in
INTEGRATION.md.  There is no PyTorch / CPU fallback: CPU tensors or a missing library raise.

Differences, all deliberate:
* one kernel pass over all frames x channels instead of ~40 ATen launches per frame; targets are
  read as the 1-byte bool they already are (the reference materialises ``.float()``, losses.py:125);
* the "no valid channel" condition (losses.py:153-161) is detected on the device and raised as the
  same ``ValueError("No valid masks")`` after ONE device->host read of T ints per call (the
  reference synchronises once per frame); ``check_valid=False`` defers/omits that read;
* M > 1 masks per channel: the reference's valid filter ``src_masks[valid]`` with a [N, M] mask flattens the masks
  into the batch dimension (losses.py:149-166), so every (channel, mask) pair is scored as its own channel with
  ``num_objects = Nv * M`` and the argmin branch (losses.py:217-229) can never be taken -- reproduced exactly that way
  (the M masks become C * M channels of the same kernel; the training wrapper itself always produces M = 1,
  sam2model.py:472-476);
* ``check_valid="deferred"``: the same ``ValueError("No valid masks")`` contract without a host synchronisation per
  call -- the minimum valid-channel count is folded into a device scalar and ``raise_if_invalid()`` reads it once
  (e.g. together with the step's loss read-back).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn

from . import _lib
from . import ops as _ops

CORE_LOSS_KEY = "total_loss"  # losses.py:17

_MODE_MULTISTEP = 0
_MODE_BCE = 1


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.Sam2B200Error(f"{what} must be a CUDA tensor: the B200 path has no CPU fallback")


def _prep_logits(x: torch.Tensor) -> torch.Tensor:
    if x.dim() == 4 and x.shape[1] == 1:
        x = x[:, 0]
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.contiguous().float()
    return x


_TICKETS_ZEROED = 0x100  # SAM2B200_LOSS_TICKETS_ZEROED (include/sam2_b200.h)
_WORKSPACES: Dict[tuple, torch.Tensor] = {}


def _persistent_workspace(dev, stream_ptr: int, nbytes: int) -> torch.Tensor:
    """Zero-filled once, then owned by (device, stream): the kernel leaves its ticket counters zeroed, so the
    forward needs no memset node (see sam2b200_mask_loss_fwd)."""
    key = (dev.index, stream_ptr)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.zeros(max(nbytes, 4096) // 4 + 1, dtype=torch.float32, device=dev)
        _WORKSPACES[key] = ws
    return ws


class _FusedMaskLossFn(torch.autograd.Function):
    """losses[4] = fused(iou_pred[T, C], logits_0 .. logits_{T-1}); see sam2b200_mask_loss_fwd."""

    @staticmethod
    def forward(ctx, cfg, targets_u8, pos_weight, iou_pred, *logits):
        # targets_u8: one uint8 tensor [T, C, H, W] or a LIST of them (the clips of a step, sum of their lengths = T): the frames of
        # all clips go through one launch with one target pointer per frame (sam2b200_mask_loss_fwd_frames)
        lib = _lib.load()
        tlist = list(targets_u8) if isinstance(targets_u8, (list, tuple)) else None
        t = len(logits)
        c, hw = logits[0].shape[0], logits[0][0].numel()
        dev = logits[0].device
        mode = cfg["mode"]
        ws_bytes = lib.sam2b200_mask_loss_workspace_bytes(t, c, hw)
        persistent = not torch.cuda.is_current_stream_capturing()
        ws = (_persistent_workspace(dev, _stream_ptr(dev), ws_bytes) if persistent
              else torch.empty(max(ws_bytes, 4) // 4, dtype=torch.float32, device=dev))
        chan_sums = torch.empty(t, c, 6, dtype=torch.float32, device=dev)
        n_valid = torch.empty(t, dtype=torch.int32, device=dev)
        losses = torch.zeros(4, dtype=torch.float32, device=dev)
        ptrs = _lib.ptr_array([x.data_ptr() for x in logits])
        if tlist is not None:
            tptr = [tt_[f].data_ptr() for tt_ in tlist for f in range(tt_.shape[0])]
            assert len(tptr) == t and all(tt_.is_contiguous() and tt_.dtype == torch.uint8 for tt_ in tlist)
            fwd_fn, targ = lib.sam2b200_mask_loss_fwd_frames, _lib.ptr_array(tptr)
        else:
            fwd_fn, targ = lib.sam2b200_mask_loss_fwd, targets_u8.data_ptr()
        with _ops._Timed("mask_loss_fwd", 5.0 * t * c * hw):
            rc = fwd_fn(
                ptrs, targ, iou_pred.data_ptr() if iou_pred is not None else None,
                pos_weight.data_ptr() if pos_weight is not None else None, ws.data_ptr(),
                chan_sums.data_ptr(), n_valid.data_ptr(), losses.data_ptr(), t, c, hw,
                mode | (_TICKETS_ZEROED if persistent else 0),
                cfg["alpha"], cfg["gamma"], cfg["inv_temp"], int(cfg["iou_l1"]), int(cfg["reduction_mean"]),
                _stream_ptr(dev))
        if rc != 0 and persistent:
            # an aborted launch may leave the ticket counters non-zero: the cached workspace is dropped (re-zeroed on next use)
            _WORKSPACES.pop((dev.index, _stream_ptr(dev)), None)
        _lib.check(rc, "sam2b200_mask_loss_fwd")
        ctx.cfg = cfg
        ctx.shape = (t, c, hw)
        ctx.logit_shapes = [tuple(x.shape) for x in logits]
        ctx.n_target_tensors = len(tlist) if tlist is not None else 0
        ctx.save_for_backward(*(tlist if tlist is not None else [targets_u8]), pos_weight, iou_pred, chan_sums, n_valid, *logits)
        ctx.mark_non_differentiable(chan_sums, n_valid)
        return losses, chan_sums, n_valid

    @staticmethod
    def backward(ctx, g_losses, _g_sums, _g_nv):
        lib = _lib.load()
        nt = max(ctx.n_target_tensors, 1)
        tlist = list(ctx.saved_tensors[:nt])
        pos_weight, iou_pred, chan_sums, n_valid, *logits = ctx.saved_tensors[nt:]
        t, c, hw = ctx.shape
        cfg = ctx.cfg
        dev = logits[0].device
        g = g_losses.contiguous().float()
        dl = torch.empty(t, c, hw, dtype=torch.float32, device=dev)
        diou = torch.empty(t, c, dtype=torch.float32, device=dev) if iou_pred is not None else None
        lp = _lib.ptr_array([x.data_ptr() for x in logits])
        dp = _lib.ptr_array([dl[f].data_ptr() for f in range(t)])
        if ctx.n_target_tensors:
            bwd_fn = lib.sam2b200_mask_loss_bwd_frames
            targ = _lib.ptr_array([tt_[f].data_ptr() for tt_ in tlist for f in range(tt_.shape[0])])
        else:
            bwd_fn, targ = lib.sam2b200_mask_loss_bwd, tlist[0].data_ptr()
        with _ops._Timed("mask_loss_bwd", 9.0 * t * c * hw):
            rc = bwd_fn(
                lp, dp, targ, iou_pred.data_ptr() if iou_pred is not None else None,
                pos_weight.data_ptr() if pos_weight is not None else None, chan_sums.data_ptr(),
                n_valid.data_ptr(), g.data_ptr(), diou.data_ptr() if diou is not None else None, t, c, hw,
                cfg["mode"], cfg["alpha"], cfg["gamma"], cfg["inv_temp"], int(cfg["iou_l1"]),
                int(cfg["reduction_mean"]), _stream_ptr(dev))
        _lib.check(rc, "sam2b200_mask_loss_bwd")
        grads = [dl[f].view(ctx.logit_shapes[f]) for f in range(t)]
        return (None, None, None, diou, *grads)


def _targets_u8(targets_batch: torch.Tensor) -> torch.Tensor:
    tb = targets_batch
    if tb.dtype == torch.bool:
        return tb.contiguous().view(torch.uint8)
    if tb.dtype == torch.uint8:
        return tb.contiguous()
    return (tb > 0).contiguous().view(torch.uint8)  # float {0,1} masks: foreground = > 0 (losses.py:64)


def reset_workspaces() -> None:
    """Drop the cached per-(device, stream) loss workspaces (they are re-created zero-filled).  Call after a sticky CUDA
    error or an interrupted launch: the forward relies on its ticket counters having been left at zero."""
    _WORKSPACES.clear()


def _raise_if_no_valid(n_valid: torch.Tensor):
    try:
        bad = bool((n_valid == 0).any().item())
    except RuntimeError:
        reset_workspaces()      # the read-back surfaced an asynchronous failure of the launch
        raise
    if bad:
        raise ValueError("No valid masks")  # losses.py:161


class MultiStepMultiMasksAndIous(nn.Module):
    """Fused focal + dice + IoU-regression loss; mirrors losses.py:79-248 of the reference."""

    def __init__(self, weight_dict, focal_alpha=0.25, focal_gamma=2.0, supervise_all_iou=False,
                 iou_use_l1_loss=False, pred_obj_scores=False, focal_gamma_obj_score=0.0,
                 focal_alpha_obj_score=-1, logit_temperature: float = 1.0, check_valid: bool = True):
        super().__init__()
        self.weight_dict = weight_dict
        self.focal_alpha = focal_alpha
        self.focal_gamma = focal_gamma
        assert "loss_mask" in self.weight_dict  # losses.py:96-98
        assert "loss_dice" in self.weight_dict
        assert "loss_iou" in self.weight_dict
        if "loss_class" not in self.weight_dict:
            self.weight_dict["loss_class"] = 0.0
        self.focal_alpha_obj_score = focal_alpha_obj_score
        self.focal_gamma_obj_score = focal_gamma_obj_score
        self.supervise_all_iou = supervise_all_iou
        self.iou_use_l1_loss = iou_use_l1_loss
        self.pred_obj_scores = pred_obj_scores
        if not (isinstance(logit_temperature, (int, float)) and logit_temperature > 0):
            raise ValueError("logit_temperature must be a positive float")  # losses.py:107-108
        self.logit_temperature = float(logit_temperature)
        self.check_valid = check_valid          # True (per call, like the reference) | False | "deferred"
        self._min_valid: Optional[torch.Tensor] = None

    def deferred_state(self) -> Optional[torch.Tensor]:
        """check_valid="deferred": int32 device scalar = minimum number of valid channels over every frame seen since the
        last raise_if_invalid() (None before the first call).  0 means a frame had no foreground."""
        return self._min_valid

    def raise_if_invalid(self, value: Optional[int] = None) -> None:
        """Deferred form of losses.py:153-161.  `value`: the content of deferred_state() if the caller already copied it
        to the host (e.g. in the same read-back as the loss); otherwise it is read here (one 4-byte D2H)."""
        if self._min_valid is None:
            return
        v = int(self._min_valid.item()) if value is None else int(value)
        self._min_valid = None
        if v <= 0:
            raise ValueError("No valid masks")  # losses.py:161

    def forward_clips(self, clips) -> Dict[str, torch.Tensor]:
        """The criterion over SEVERAL clips of a step in one launch: `clips` = [(outs_batch, targets_batch), ...], one pair per
        clip exactly as `forward` takes them (the reference trainer calls the criterion once per batch element,
        trainer.py:268,303).  Returns the SUM over the clips of the dict `forward` returns for each -- what a trainer that
        averages the batch back-propagates, up to its 1 / len(clips) -- from one kernel launch over all frames of all clips (one
        target pointer per frame; no concatenation of the targets).  Same validity contract: a frame without any valid channel in
        ANY clip raises "No valid masks" (per call, or deferred)."""
        if len(clips) == 0:
            raise ValueError("forward_clips needs at least one clip")
        outs_all, targets_all = [], []
        for outs_batch, targets_batch in clips:
            assert len(outs_batch) == len(targets_batch)  # losses.py:113
            outs_all += list(outs_batch)
            targets_all.append(targets_batch)
        shapes = {tuple(tb.shape[1:]) for tb in targets_all}
        if len(shapes) != 1:
            raise ValueError("forward_clips: every clip must have the same [C, H, W]")
        return self._forward_frames(outs_all, targets_all)

    def forward(self, outs_batch: List[Dict], targets_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        assert len(outs_batch) == len(targets_batch)  # losses.py:113
        return self._forward_frames(outs_batch, [targets_batch])

    def _forward_frames(self, outs_batch: List[Dict], targets_list: List[torch.Tensor]) -> Dict[str, torch.Tensor]:
        multi = len(targets_list) > 1
        targets_batch = targets_list[0]          # shape reference ([C, H, W] is the same in every clip)
        for tb in targets_list:
            _require_cuda(tb, "targets_batch")
        t = len(outs_batch)
        n_steps = None
        for outs in outs_batch:
            a, b, c_ = (outs["multistep_pred_multimasks_high_res"], outs["multistep_pred_ious"],
                        outs["multistep_object_score_logits"])
            assert len(a) == len(b)  # losses.py:130-131
            assert len(c_) == len(b)
            if n_steps is None:
                n_steps = len(a)
            elif n_steps != len(a):
                raise NotImplementedError("frames with different numbers of correction steps")
        n_masks = int(outs_batch[0]["multistep_pred_multimasks_high_res"][0].shape[1]) if t else 1
        if n_masks > 1:
            # [C, M, H, W]: the reference's filter turns every (channel, mask) pair into its own row (module docstring)
            if self.pred_obj_scores:
                raise NotImplementedError("pred_obj_scores with M > 1 masks per channel (the reference itself fails to index "
                                          "object_score_logits with its [N, M] valid mask, losses.py:169-170)")
            targets_list = [tb.repeat_interleave(n_masks, dim=1) for tb in targets_list]
            targets_batch = targets_list[0]
        tu8 = [_targets_u8(tb) for tb in targets_list] if multi else _targets_u8(targets_batch)
        cfg = dict(mode=_MODE_MULTISTEP, alpha=float(self.focal_alpha), gamma=float(self.focal_gamma),
                   inv_temp=1.0 / self.logit_temperature, iou_l1=bool(self.iou_use_l1_loss),
                   reduction_mean=True)
        total4 = None
        loss_class = None
        for s in range(n_steps):
            logits = []
            for outs in outs_batch:
                x = outs["multistep_pred_multimasks_high_res"][s]
                _require_cuda(x, "mask logits")
                if x.dim() != 4 or x.shape[1] != n_masks:
                    raise ValueError(f"mask logits must be [C, {n_masks}, H, W] in every frame and step, got {tuple(x.shape)}")
                if n_masks > 1:
                    x = x.reshape(x.shape[0] * n_masks, 1, *x.shape[-2:])
                if tuple(x.shape[-2:]) != tuple(targets_batch.shape[-2:]) or x.shape[0] != targets_batch.shape[1]:
                    raise ValueError("mask logits / targets shape mismatch")
                logits.append(_prep_logits(x))
            ious = torch.stack([outs["multistep_pred_ious"][s].reshape(-1) for outs in outs_batch]).float()
            losses4, chan_sums, n_valid = _FusedMaskLossFn.apply(cfg, tu8, None, ious.contiguous(), *logits)
            if self.check_valid == "deferred":
                mv = n_valid.amin()
                self._min_valid = mv if self._min_valid is None else torch.minimum(self._min_valid, mv)
            elif self.check_valid:
                _raise_if_no_valid(n_valid)
            total4 = losses4 if total4 is None else total4 + losses4
            if self.pred_obj_scores:  # losses.py:194-204 -- [C, 1] tensors, not on the hot path
                valid = (chan_sums[..., 3] > 0).float()  # [T, C]; target_obj == 1 on every valid channel
                nv = valid.sum(-1, keepdim=True).clamp(min=1.0)
                osl = torch.stack([outs["multistep_object_score_logits"][s].reshape(-1) for outs in outs_batch]).float()
                ce = torch.nn.functional.softplus(-osl)  # BCE with target 1
                p = torch.sigmoid(osl)
                fl = ce * (1 - p) ** self.focal_gamma_obj_score
                if self.focal_alpha_obj_score >= 0:
                    fl = self.focal_alpha_obj_score * fl
                lc = (fl * valid / nv).sum()
                loss_class = lc if loss_class is None else loss_class + lc
        losses = {"loss_mask": total4[0], "loss_dice": total4[1], "loss_iou": total4[2],
                  "loss_class": loss_class if loss_class is not None else total4[3]}
        losses[CORE_LOSS_KEY] = self.reduce_loss(losses)
        return losses

    def reduce_loss(self, losses):  # losses.py:240-248
        reduced_loss = 0.0
        for loss_key, weight in self.weight_dict.items():
            if loss_key not in losses:
                raise ValueError(f"{type(self)} doesn't compute {loss_key}")
            if weight != 0:
                reduced_loss = reduced_loss + losses[loss_key] * weight
        return reduced_loss


class BCECategoryLoss(nn.Module):
    """Fused per-category BCE-with-logits; mirrors losses.py:251-372 of the reference."""

    def __init__(self, pos_weight: Optional[Union[List[float], torch.Tensor]] = None,
                 reduction: str = "mean", logit_temperature: float = 1.0):
        super().__init__()
        if isinstance(pos_weight, list):
            self.register_buffer("_pos_weight", torch.tensor(pos_weight, dtype=torch.float32), persistent=False)
        elif isinstance(pos_weight, torch.Tensor):
            self.register_buffer("_pos_weight", pos_weight.to(dtype=torch.float32), persistent=False)
        else:
            self._pos_weight = None  # type: ignore
        self.reduction = reduction
        if not (isinstance(logit_temperature, (int, float)) and logit_temperature > 0):
            raise ValueError("logit_temperature must be a positive float")
        self.logit_temperature = float(logit_temperature)

    def forward(self, outs_batch: List[Dict], targets_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        assert len(outs_batch) == len(targets_batch), (
            f"Mismatched sequence lengths: outs={len(outs_batch)} vs targets={len(targets_batch)}")
        if self.reduction not in ("mean", "sum"):
            raise NotImplementedError("BCECategoryLoss on the B200 path supports reduction='mean'|'sum'")
        _require_cuda(targets_batch, "targets_batch")
        logits = []
        for outs, targets in zip(outs_batch, targets_batch):
            x = outs.get("pred_masks_high_res")
            if x is None:
                x = outs.get("pred_masks")
            if x is None:
                raise KeyError("BCECategoryLoss expects 'pred_masks_high_res' or 'pred_masks' in outputs")
            if not ((x.dim() == 4 and x.shape[1] == 1) or x.dim() == 3):
                raise ValueError(f"Unexpected logits shape for BCECategoryLoss: {tuple(x.shape)}")
            if targets.dim() != 3:
                raise ValueError(f"Unexpected target shape for BCECategoryLoss: {tuple(targets.shape)}")
            _require_cuda(x, "mask logits")
            logits.append(_prep_logits(x))
        c = logits[0].shape[0]
        pw = None
        if self._pos_weight is not None:
            pw = self._pos_weight.to(device=logits[0].device, dtype=torch.float32).reshape(-1).contiguous()
        tu8 = _targets_u8(targets_batch)
        cfg = dict(mode=_MODE_BCE, alpha=0.0, gamma=0.0, inv_temp=1.0 / self.logit_temperature, iou_l1=False,
                   reduction_mean=self.reduction == "mean")
        losses4, chan_sums, n_valid = _FusedMaskLossFn.apply(cfg, tu8, pw, None, *logits)
        if pw is not None:
            # the reference compares len(pos_weight) with the number of VALID channels (losses.py:359-362)
            if pw.numel() != c or bool((n_valid != c).any().item()):
                raise ValueError(f"pos_weight length {pw.numel()} does not match number of classes")
        total_loss = losses4[0] / max(len(outs_batch), 1)  # losses.py:368
        return {"loss_bce": total_loss, CORE_LOSS_KEY: total_loss}


# ---- functional forms (losses.py:20-76): same names, signatures and results, fused kernels underneath ----------
class _ChannelSumsFn(torch.autograd.Function):
    """sums[C, 6] = (sum focal, sum p*t, sum p, sum t, |pred & gt|, |pred | gt|) per channel of x [C, HW] against the
    0/1 targets t [C, HW]; differentiable in x through the first three sums (sam2b200_mask_loss_bwd_coef)."""

    @staticmethod
    def forward(ctx, x, t_u8, alpha, gamma):
        lib = _lib.load()
        c, hw = x.shape
        dev = x.device
        ws_bytes = lib.sam2b200_mask_loss_workspace_bytes(1, c, hw)
        ws = torch.empty(max(ws_bytes, 4) // 4, dtype=torch.float32, device=dev)
        sums = torch.empty(1, c, 6, dtype=torch.float32, device=dev)
        n_valid = torch.empty(1, dtype=torch.int32, device=dev)
        losses = torch.zeros(4, dtype=torch.float32, device=dev)
        iou_dummy = torch.zeros(1, c, dtype=torch.float32, device=dev)
        rc = lib.sam2b200_mask_loss_fwd(_lib.ptr_array([x.data_ptr()]), t_u8.data_ptr(), iou_dummy.data_ptr(), None,
                                        ws.data_ptr(), sums.data_ptr(), n_valid.data_ptr(), losses.data_ptr(), 1, c, hw,
                                        _MODE_MULTISTEP, float(alpha), float(gamma), 1.0, 0, 1, _stream_ptr(dev))
        _lib.check(rc, "sam2b200_mask_loss_fwd")
        ctx.save_for_backward(x, t_u8)
        ctx.cfg = (float(alpha), float(gamma))
        return sums[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, t_u8 = ctx.saved_tensors
        c, hw = x.shape
        g = g.float()
        coef = torch.stack([g[:, 0], g[:, 1] + g[:, 2], g[:, 2]], dim=1).contiguous()
        dx = torch.empty_like(x)
        rc = lib.sam2b200_mask_loss_bwd_coef(_lib.ptr_array([x.data_ptr()]), _lib.ptr_array([dx.data_ptr()]), t_u8.data_ptr(),
                                             coef.data_ptr(), 1, c, hw, ctx.cfg[0], ctx.cfg[1], 1.0, _stream_ptr(x.device))
        _lib.check(rc, "sam2b200_mask_loss_bwd_coef")
        return dx, None, None, None


def _channel_sums(inputs: torch.Tensor, targets: torch.Tensor, lead_dims: int, alpha: float = -1.0, gamma: float = 0.0):
    """inputs / targets flattened to [prod(shape[:lead_dims]), rest] -> sums [.., 6] with the leading shape restored."""
    _require_cuda(inputs, "inputs")
    _require_cuda(targets, "targets")
    lead = tuple(inputs.shape[:lead_dims])
    x = inputs.reshape(int(torch.Size(lead).numel()), -1).float().contiguous()
    t = targets.reshape(x.shape[0], -1)
    if t.shape != x.shape:
        raise ValueError("inputs / targets shape mismatch")
    if t.dtype not in (torch.bool, torch.uint8) and bool(((t != 0) & (t != 1)).any().item()):
        raise NotImplementedError("the B200 loss kernels take binary {0, 1} targets")
    t_u8 = (t != 0).contiguous().view(torch.uint8) if t.dtype != torch.uint8 else t.contiguous()
    return _ChannelSumsFn.apply(x, t_u8, alpha, gamma).view(*lead, 6), x.shape[1]


def dice_loss(inputs, targets, num_objects, loss_on_multimask=False):
    """losses.py:20-34.  One fused pass over the logits instead of four."""
    if loss_on_multimask:
        assert inputs.dim() == 4 and targets.dim() == 4
        s, _ = _channel_sums(inputs, targets, 2)
    else:
        s, _ = _channel_sums(inputs, targets, 1)
    loss = 1 - (2 * s[..., 1] + 1) / (s[..., 2] + s[..., 3] + 1)
    if loss_on_multimask:
        return loss / num_objects
    return loss.sum() / num_objects


def sigmoid_focal_loss(inputs, targets, num_objects, alpha: float = 0.25, gamma: float = 2, loss_on_multimask=False):
    """losses.py:37-57."""
    if loss_on_multimask:
        assert inputs.dim() == 4
        s, hw = _channel_sums(inputs, targets, 2, alpha, gamma)
        return s[..., 0] / hw / num_objects                       # flatten(2).mean(-1) / num_objects
    # loss.mean(1).sum() / num_objects: the mean runs over dim 1 only, everything else is summed
    s, _ = _channel_sums(inputs, targets, 1, alpha, gamma)
    return s[..., 0].sum() / inputs.shape[1] / num_objects


def iou_loss(inputs, targets, pred_ious, num_objects, loss_on_multimask=False, use_l1_loss=False):
    """losses.py:60-76.  Differentiable in pred_ious only (the mask comparison has no gradient)."""
    assert inputs.dim() == 4 and targets.dim() == 4
    with torch.no_grad():
        s, _ = _channel_sums(inputs.detach(), (targets > 0), 2)
        actual_ious = s[..., 4] / torch.clamp(s[..., 5], min=1.0)
    if use_l1_loss:
        loss = torch.nn.functional.l1_loss(pred_ious, actual_ious, reduction="none")
    else:
        loss = torch.nn.functional.mse_loss(pred_ious, actual_ious, reduction="none")
    if loss_on_multimask:
        return loss / num_objects
    return loss.sum() / num_objects
