"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU (fp32).

Run in the build container only (needs /root/reference):  ``python -m oracle.make_golden``.
The fixtures pin the oracle (tests/test_oracle_golden.py) and travel to the GPU box, where the
reference itself does not exist.  Inputs are analytic (oracle/detgen.py), so only outputs are
stored.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import detgen, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_attention(ns, grid, batch, n_frames, n_ptr, tag, full_grads=False):
    torch.manual_seed(0)
    model = ref_shim.build_memory_attention(ns).eval()  # eval: dropout off (parity contract)
    params = detgen.det_params([(n, tuple(p.shape)) for n, p in model.named_parameters()])
    assert [n for n, _ in model.named_parameters()] == [n for n, _ in detgen.param_shapes()]
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(params[n])
    inp = detgen.attention_inputs(grid, batch, n_frames, n_ptr)
    curr = inp["curr"].clone().requires_grad_(True)
    curr_pos = inp["curr_pos"].clone().requires_grad_(True)
    memory = inp["memory"].clone().requires_grad_(True)
    memory_pos = inp["memory_pos"].clone().requires_grad_(True)
    out = model(curr=[curr], curr_pos=[curr_pos], memory=memory, memory_pos=memory_pos,
                num_obj_ptr_tokens=n_ptr)
    out.backward(inp["grad_out"])
    rec = dict(
        grid=grid, batch=batch, n_frames=n_frames, n_ptr=n_ptr,
        out=out.detach().numpy(),
        d_curr=curr.grad.numpy(), d_curr_pos=curr_pos.grad.numpy(),
        d_memory=memory.grad.numpy(), d_memory_pos=memory_pos.grad.numpy(),
    )
    names, sums = [], []
    for n, p in model.named_parameters():
        names.append(n)
        sums.append(float(p.grad.abs().sum()))
        if full_grads or n in ("layers.0.self_attn.q_proj.weight", "layers.3.cross_attn_image.k_proj.weight",
                               "layers.1.cross_attn_image.v_proj.bias", "layers.2.linear1.bias",
                               "layers.0.norm2.weight", "norm.bias"):
            rec["dparam:" + n] = p.grad.numpy()
    rec["param_names"] = np.array(names)
    rec["param_grad_abs_sums"] = np.array(sums, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, f"attn_{tag}.npz"), **rec)
    return rec


REFINIT_PARAM_GRADS = ("layers.0.self_attn.q_proj.weight", "layers.3.cross_attn_image.k_proj.weight",
                       "layers.1.cross_attn_image.v_proj.bias", "layers.2.linear1.bias", "layers.0.norm2.weight", "norm.bias",
                       "layers.3.cross_attn_image.out_proj.weight", "layers.1.self_attn.v_proj.weight")


def golden_attention_refinit(ns, grid, batch, n_frames, n_ptr, tag, seed=1234):
    """WELL-CONDITIONED fixture: the reference stack with ITS OWN random init (torch.manual_seed(0), nn.Linear defaults,
    get_clones) on N(0,1) inputs (attention_oracle.random_inputs).  Stored: output, all input gradients, eight full
    parameter gradients, per-parameter |grad| sums, and checksums of weights / inputs so that a consumer which
    regenerates them (attention_oracle.reference_init_params / random_inputs) can prove it holds the same tensors."""
    from . import attention_oracle as ao
    torch.manual_seed(0)
    model = ref_shim.build_memory_attention(ns).eval()
    mine = ao.reference_init_params(0)
    for n, p in model.named_parameters():       # the regenerated init IS the reference's
        assert torch.equal(p.detach(), mine[n]), n
    inp = ao.random_inputs(grid, batch, n_frames, n_ptr, seed)
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("curr", "curr_pos", "memory", "memory_pos")}
    out = model(curr=[leaves["curr"]], curr_pos=[leaves["curr_pos"]], memory=leaves["memory"],
                memory_pos=leaves["memory_pos"], num_obj_ptr_tokens=n_ptr)
    out.backward(inp["grad_out"])
    rec = dict(grid=grid, batch=batch, n_frames=n_frames, n_ptr=n_ptr, seed=seed, out=out.detach().numpy(),
               d_curr=leaves["curr"].grad.numpy(), d_memory=leaves["memory"].grad.numpy(),
               d_memory_pos=leaves["memory_pos"].grad.numpy(),
               d_curr_pos_over_d_curr=float((leaves["curr_pos"].grad.double() * leaves["curr"].grad.double()).sum()
                                            / (leaves["curr"].grad.double() ** 2).sum()),
               input_abs_sums=np.array([float(inp[k].double().abs().sum()) for k in ("curr", "curr_pos", "memory", "memory_pos", "grad_out")]),
               weight_abs_sums=np.array([float(p.detach().double().abs().sum()) for _, p in model.named_parameters()]),
               param_names=np.array([n for n, _ in model.named_parameters()]),
               param_grad_abs_sums=np.array([float(p.grad.double().abs().sum()) for _, p in model.named_parameters()]))
    for n, p in model.named_parameters():
        if n in REFINIT_PARAM_GRADS:
            rec["dparam:" + n] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"attn_refinit_{tag}.npz"), **rec)
    return rec


def golden_bank():
    """Memory-bank assembly: the unmodified SAM2Base._prepare_memory_conditioned_features (sam2_base.py:524-713) with a
    stub memory_attention that records (memory, memory_pos, num_obj_ptr_tokens), plus the gradients that reach the
    trainable maskmem_tpos_enc and obj_ptr_tpos_proj through memory_pos."""
    import types
    SAM2Base = ref_shim.load_sam2_base()

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.out_proj = torch.nn.Conv2d(256, 64, 1)

    class Img(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.neck = types.SimpleNamespace(d_model=256)

    class Rec(torch.nn.Module):
        d_model = 256

        def forward(self, curr, curr_pos, memory, memory_pos, num_obj_ptr_tokens):
            self.rec = (memory, memory_pos, num_obj_ptr_tokens)
            return curr[0]

    for tag, kw in detgen.bank_scenarios():
        h = w = 6
        base = SAM2Base(image_encoder=Img(), memory_attention=Rec(), memory_encoder=Enc(), image_size=96, backbone_stride=16,
                        num_maskmem=7, use_obj_ptrs_in_encoder=True, max_obj_ptrs_in_encoder=16, add_tpos_enc_to_obj_ptrs=True,
                        proj_tpos_enc_in_obj_ptrs=True, use_signed_tpos_enc_to_obj_ptrs=True,
                        only_obj_ptrs_in_the_past_for_eval=True, directly_add_no_mem_embed=True,
                        max_cond_frames_in_attn=kw.get("max_cond", -1), memory_temporal_stride_for_eval=kw.get("stride", 1))
        od, tpos, pw, pb = detgen.bank_inputs(kw["cond"], kw["non_cond"], h=h, w=w)
        with torch.no_grad():
            base.maskmem_tpos_enc.copy_(tpos)
            base.obj_ptr_tpos_proj.weight.copy_(pw)
            base.obj_ptr_tpos_proj.bias.copy_(pb)
        base.train(kw["training"])
        feats = [detgen.det((h * w, 2, 256), 0.31, 0.7, 1.0)]
        base._prepare_memory_conditioned_features(kw["frame_idx"], False, feats, [feats[0] * 0.5], [(h, w)], od, kw["num_frames"],
                                                  track_in_reverse=kw.get("reverse", False))
        memory, memory_pos, n_ptr = base.memory_attention.rec
        wgt = detgen.det(tuple(memory_pos.shape), 0.013, 0.9, 1.0)
        (memory_pos * wgt).sum().backward()
        np.savez_compressed(os.path.join(OUT, f"bank_{tag}.npz"), memory=memory.detach().numpy(), memory_pos=memory_pos.detach().numpy(),
                            n_ptr=n_ptr, d_tpos=base.maskmem_tpos_enc.grad.numpy(), d_proj_w=base.obj_ptr_tpos_proj.weight.grad.numpy(),
                            d_proj_b=base.obj_ptr_tpos_proj.bias.grad.numpy())
        print("bank", tag, tuple(memory.shape), n_ptr)


def golden_functional(ns, n, m, s, tag):
    """Stand-alone dice_loss / sigmoid_focal_loss / iou_loss of the reference (losses.py:20-76), both branches."""
    L = ns.losses
    logits, targets, iou_pred = detgen.loss_inputs(n, m, s)          # [n, m, 1, s, s], bool [n, m, s, s], [n, m, 1]
    x = logits[:, :, 0].clone().requires_grad_(True)                # [n, m, s, s]
    t = targets.float()
    iou = iou_pred[:, :, 0].clone().requires_grad_(True)            # [n, m]
    rec = {}
    outs = {"mm:dice": L.dice_loss(x, t, 3.0, loss_on_multimask=True),
            "mm:focal": L.sigmoid_focal_loss(x, t, 3.0, alpha=0.25, gamma=2, loss_on_multimask=True),
            "mm:iou": L.iou_loss(x, t, iou, 3.0, loss_on_multimask=True, use_l1_loss=False),
            "flat:dice": L.dice_loss(x.flatten(1), t.flatten(1), 3.0),
            "flat:focal": L.sigmoid_focal_loss(x.flatten(1), t.flatten(1), 3.0, alpha=-1, gamma=0),
            "flat:iou": L.iou_loss(x, t, iou, 3.0, use_l1_loss=True)}
    tot = 0
    for i, (k, v) in enumerate(outs.items()):
        rec[k] = v.detach().numpy()
        w = detgen.det(tuple(v.shape), 0.77 + i, 0.3, 1.0) + 1.5 if v.dim() else torch.tensor(1.0 + 0.25 * i)
        rec[k + ":w"] = w.numpy()
        tot = tot + (v * w).sum()
    tot.backward()
    rec["dx"] = x.grad.numpy()
    rec["diou"] = iou.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"lossfn_{tag}.npz"), **rec)
    return rec


def golden_merged(ns, t, n_obj, c, s, tag):
    """Producer side of the loss through the UNMODIFIED reference: F.interpolate exactly as sam2_base.py:393-399 calls it,
    merge_object_results_to_category (utils/masks.py:53-212) and MultiStepMultiMasksAndIous."""
    import torch.nn.functional as F
    masks = ref_shim.load_masks()
    low, iou_pred, obj_to_cat, targets = detgen.merged_inputs(t, n_obj, c, s)
    rec = dict(t=t, n_obj=n_obj, c=c, s=s, obj_to_cat=np.asarray(obj_to_cat))
    for mode, kw in (("l1", dict(iou_use_l1_loss=True)), ("mse", dict(iou_use_l1_loss=False)),
                     ("temp", dict(iou_use_l1_loss=True, logit_temperature=1.6, focal_alpha=0.6))):
        crit = ns.MultiStepMultiMasksAndIous(
            weight_dict={"loss_mask": 20, "loss_dice": 1, "loss_iou": 1, "loss_class": 0},
            supervise_all_iou=True, pred_obj_scores=False, focal_gamma_obj_score=0.0, focal_alpha_obj_score=-1.0, **kw)
        x = low.clone().requires_grad_(True)
        ip = iou_pred.clone().requires_grad_(True)
        stages = []
        for f in range(t):
            hi = F.interpolate(x[f].float(), size=(4 * s, 4 * s), mode="bilinear", align_corners=False)
            stages.append({"pred_masks_high_res": hi, "multistep_pred_multimasks_high_res": [hi],
                           "multistep_pred_ious": [ip[f]], "multistep_object_score_logits": [torch.zeros(n_obj, 1)],
                           "point_inputs": None, "mask_inputs": None})
        merged = masks.merge_object_results_to_category(stages, obj_to_cat, c)
        if mode == "l1":
            rec["merged_logits"] = torch.stack([m["multistep_pred_multimasks_high_res"][0] for m in merged]).detach().numpy()
            rec["merged_ious"] = torch.stack([m["multistep_pred_ious"][0] for m in merged]).detach().numpy()
        losses = crit(merged, targets)
        losses["total_loss"].backward()
        for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
            rec[f"{mode}:{k}"] = float(losses[k])
        rec[f"{mode}:dlow"] = x.grad.numpy()
        rec[f"{mode}:diou"] = ip.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"merged_{tag}.npz"), **rec)
    return rec


def golden_loss(ns, t, c, s, tag):
    logits, targets, iou_pred = detgen.loss_inputs(t, c, s)
    rec = dict(t=t, c=c, s=s)
    for mode, l1 in (("l1", True), ("mse", False)):
        crit = ns.MultiStepMultiMasksAndIous(
            weight_dict={"loss_mask": 20, "loss_dice": 1, "loss_iou": 1, "loss_class": 0},
            supervise_all_iou=True, iou_use_l1_loss=l1, pred_obj_scores=False,
            focal_gamma_obj_score=0.0, focal_alpha_obj_score=-1.0)
        x = logits.clone().requires_grad_(True)
        ip = iou_pred.clone().requires_grad_(True)
        outs = [{"multistep_pred_multimasks_high_res": [x[f]], "multistep_pred_ious": [ip[f]],
                 "multistep_object_score_logits": [torch.zeros(c, 1)]} for f in range(t)]
        losses = crit(outs, targets)
        losses["total_loss"].backward()
        for k in ("loss_mask", "loss_dice", "loss_iou", "loss_class", "total_loss"):
            rec[f"{mode}:{k}"] = float(losses[k].detach()) if torch.is_tensor(losses[k]) else float(losses[k])
        rec[f"{mode}:dlogits"] = x.grad.numpy()
        rec[f"{mode}:diou"] = ip.grad.numpy()
    # temperature variant
    crit = ns.MultiStepMultiMasksAndIous(
        weight_dict={"loss_mask": 1, "loss_dice": 10, "loss_iou": 10}, iou_use_l1_loss=True,
        logit_temperature=2.5, focal_alpha=0.6, focal_gamma=2.0)
    x = logits.clone().requires_grad_(True)
    ip = iou_pred.clone().requires_grad_(True)
    outs = [{"multistep_pred_multimasks_high_res": [x[f]], "multistep_pred_ious": [ip[f]],
             "multistep_object_score_logits": [torch.zeros(c, 1)]} for f in range(t)]
    losses = crit(outs, targets)
    losses["total_loss"].backward()
    for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
        rec[f"temp:{k}"] = float(losses[k])
    rec["temp:dlogits"] = x.grad.numpy()
    # BCE category loss
    for tagb, kw in (("bce", {}), ("bce_pw", dict(pos_weight=[1.5, 0.5, 2.0][:c] + [1.0] * max(0, c - 3),
                                                  logit_temperature=1.7))):
        crit = ns.BCECategoryLoss(**kw)
        x = logits.clone().requires_grad_(True)
        tg = targets.clone()
        if "pos_weight" in kw:  # the reference requires every channel valid when pos_weight is set
            tg[:, :, 0, 0] = True
        outs = [{"pred_masks_high_res": x[f]} for f in range(t)]
        losses = crit(outs, tg)
        losses["total_loss"].backward()
        rec[f"{tagb}:total_loss"] = float(losses["total_loss"])
        rec[f"{tagb}:dlogits"] = x.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"loss_{tag}.npz"), **rec)
    return rec


def golden_loss_multimask(ns, t, c, m, s, tag):
    """MultiStepMultiMasksAndIous of the unmodified reference on M > 1 masks per channel (object_score_logits None: with
    a tensor the reference fails to index it with its [N, M] valid mask, losses.py:169-170)."""
    logits, targets, iou_pred = detgen.multimask_loss_inputs(t, c, m, s)
    rec = dict(t=t, c=c, m=m, s=s)
    for mode, kw in (("l1_all", dict(iou_use_l1_loss=True, supervise_all_iou=True)), ("mse", dict(iou_use_l1_loss=False))):
        crit = ns.MultiStepMultiMasksAndIous(weight_dict={"loss_mask": 20, "loss_dice": 1, "loss_iou": 1, "loss_class": 0}, **kw)
        x = logits.clone().requires_grad_(True)
        ip = iou_pred.clone().requires_grad_(True)
        outs = [{"multistep_pred_multimasks_high_res": [x[f]], "multistep_pred_ious": [ip[f]],
                 "multistep_object_score_logits": [None]} for f in range(t)]
        losses = crit(outs, targets)
        losses["total_loss"].backward()
        for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
            rec[f"{mode}:{k}"] = float(losses[k])
        rec[f"{mode}:dlogits"] = x.grad.numpy()
        rec[f"{mode}:diou"] = ip.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"loss_multimask_{tag}.npz"), **rec)
    return rec


MEMENC_FULL_GRADS = ("mask_downsampler.encoder.0.weight", "mask_downsampler.encoder.1.weight", "mask_downsampler.encoder.4.bias",
                     "mask_downsampler.encoder.6.weight", "mask_downsampler.encoder.10.weight", "mask_downsampler.encoder.12.bias",
                     "pix_feat_proj.bias", "fuser.layers.0.dwconv.weight", "fuser.layers.0.norm.weight", "fuser.layers.1.gamma",
                     "fuser.layers.1.pwconv1.bias", "fuser.layers.0.pwconv2.bias", "out_proj.weight", "fuser.layers.1.dwconv.bias")


def golden_memory_encoder(b, grid, tag, skip_mask_sigmoid, seed=77):
    """The UNMODIFIED reference MemoryEncoder (its own init under torch.manual_seed(0); layer scales moved to 0.5 +- 0.1, see
    memenc_oracle.reference_init_state) on N(0, 1) features / N(0, 16) mask logits: outputs, input gradients, per-parameter
    |grad| sums, fourteen full parameter gradients, checksums of the regenerated weights / inputs."""
    from . import memenc_oracle as mo
    torch.manual_seed(0)
    model = ref_shim.build_memory_encoder()
    sd = mo.reference_init_state(0)
    with torch.no_grad():
        for i in range(2):
            model.fuser.layers[i].gamma.copy_(sd[f"fuser.layers.{i}.gamma"])
    ref_sd = dict(model.named_parameters())
    assert list(ref_sd.keys()) == list(dict(model.named_parameters()).keys()) and set(ref_sd) == set(sd), set(ref_sd) ^ set(sd)
    for n, p in ref_sd.items():
        assert torch.equal(p.detach(), sd[n]), n           # the regenerated init IS the reference's
    inp = mo.random_inputs(b, grid, seed)
    pix = inp["pix_feat"].clone().requires_grad_(True)
    masks = inp["masks"].clone().requires_grad_(True)
    m_in = torch.sigmoid(masks) * 20.0 - 10.0 if skip_mask_sigmoid else masks     # sam2_base.py:741-747 scales before the call
    out = model(pix, m_in, skip_mask_sigmoid=skip_mask_sigmoid)
    out["vision_features"].backward(inp["grad_out"])
    names = [n for n, _ in model.named_parameters()]
    rec = dict(b=b, grid=grid, seed=seed, skip=int(skip_mask_sigmoid), features=out["vision_features"].detach().numpy(),
               pos=out["vision_pos_enc"][0].detach().numpy(), d_pix_feat=pix.grad.numpy(), d_masks=masks.grad.numpy(),
               param_names=np.array(names), weight_abs_sums=np.array([float(ref_sd[n].detach().double().abs().sum()) for n in names]),
               param_grad_abs_sums=np.array([float(ref_sd[n].grad.double().abs().sum()) for n in names]),
               input_abs_sums=np.array([float(inp[k].double().abs().sum()) for k in ("pix_feat", "masks", "grad_out")]))
    for n in MEMENC_FULL_GRADS:
        rec["dparam:" + n] = ref_sd[n].grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"memenc_{tag}.npz"), **rec)
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.load()
    torch.set_num_threads(8)
    r = golden_attention(ns, 4, 2, 2, 8, "g4_b2_f2_p8", full_grads=False)
    print("attn g4: out sum", float(r["out"].sum()), "abs", float(np.abs(r["out"]).sum()),
          "d_curr abs", float(np.abs(r["d_curr"]).sum()))
    r = golden_attention(ns, 8, 3, 3, 12, "g8_b3_f3_p12")
    print("attn g8: out abs", float(np.abs(r["out"]).sum()))
    r = golden_attention(ns, 12, 1, 1, 0, "g12_b1_f1_p0")
    print("attn g12: out abs", float(np.abs(r["out"]).sum()))
    r = golden_attention_refinit(ns, 24, 1, 7, 28, "g24_b1_f7_p28")      # BASELINE configs[0] steady state
    print("attn refinit g24: out abs", float(np.abs(r["out"]).sum()))
    r = golden_attention_refinit(ns, 8, 3, 3, 12, "g8_b3_f3_p12", seed=4321)
    print("attn refinit g8: out abs", float(np.abs(r["out"]).sum()))
    r = golden_memory_encoder(2, 4, "b2_g4_sigmoid", False)
    print("memory encoder: |features|", float(np.abs(r["features"]).sum()))
    r = golden_memory_encoder(3, 6, "b3_g6_scaled", True, seed=78)
    print("memory encoder (scaled masks): |features|", float(np.abs(r["features"]).sum()))
    r = golden_loss(ns, 2, 3, 16, "t2_c3_s16")
    print("loss: l1 total", r["l1:total_loss"], "mse total", r["mse:total_loss"], "bce", r["bce:total_loss"])
    r = golden_loss_multimask(ns, 2, 3, 3, 16, "t2_c3_m3_s16")
    print("multimask loss: total", r["l1_all:total_loss"])
    golden_functional(ns, 3, 2, 24, "n3_m2_s24")
    golden_bank()
    r = golden_merged(ns, 2, 5, 4, 6, "t2_n5_c4_s6")
    print("merged: l1 total", r["l1:total_loss"], "mse", r["mse:total_loss"])
    r = golden_merged(ns, 2, 7, 3, 11, "t2_n7_c3_s11")
    print("merged2: l1 total", r["l1:total_loss"])
    r = golden_loss(ns, 3, 5, 40, "t3_c5_s40")
    print("loss2: l1 total", r["l1:total_loss"])


if __name__ == "__main__":
    main()
