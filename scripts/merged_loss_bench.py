"""Fused up-sample + category merge + loss (csrc/merge_loss.cu) vs the un-fused chain on the same B200:
torch F.interpolate + per-category max / weighted IoU average (ATen) feeding this package's fused mask loss.
CUDA events, rotating over `nsets` input sets so that targets (1 B/px) come from HBM, not L2."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200.merged_loss import CategoryMergedMultiStepLoss
from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous

W = {"loss_mask": 20, "loss_dice": 1, "loss_iou": 1, "loss_class": 0}
dev = torch.device("cuda:0")


def groups_of(o2c, c):
    g = [[] for _ in range(c)]
    for i, k in enumerate(o2c):
        g[k].append(i)
    return g


def unfused(low, ious, groups, tg, crit, S):
    outs = []
    for f in range(len(low)):
        hi = torch.nn.functional.interpolate(low[f].float(), size=(S, S), mode="bilinear", align_corners=False)
        w = torch.sigmoid(hi).sum(dim=(1, 2, 3))
        xs, qs = [], []
        for idx in groups:
            xs.append(hi[idx].max(dim=0).values)
            sw = w[idx].view(-1, 1)
            qs.append((ious[f][idx] * sw).sum(0) / sw.sum(0))
        outs.append({"multistep_pred_multimasks_high_res": [torch.stack(xs)], "multistep_pred_ious": [torch.stack(qs)],
                     "multistep_object_score_logits": [None]})
    return crit(outs, tg)


def bench(name, T, C, per_cat, s, nsets, iters):
    S = 4 * s
    n_obj = C * per_cat
    o2c = [i % C for i in range(n_obj)]
    groups = groups_of(o2c, C)
    g = torch.Generator().manual_seed(3)
    sets = []
    for _ in range(nsets):
        low = [(torch.randn(n_obj, 1, s, s, generator=g) * 4).to(dev).requires_grad_(True) for _ in range(T)]
        ious = [torch.rand(n_obj, 1, generator=g).to(dev).requires_grad_(True) for _ in range(T)]
        tg = (torch.rand(T, C, S, S, generator=g) > 0.8).to(dev)
        sets.append((low, ious, tg))
    fused = CategoryMergedMultiStepLoss(dict(W), supervise_all_iou=True, iou_use_l1_loss=True, check_valid=False)
    plain = MultiStepMultiMasksAndIous(dict(W), supervise_all_iou=True, iou_use_l1_loss=True, check_valid=False)

    def run_fused(k):
        low, ious, tg = sets[k % nsets]
        st = [{"multistep_pred_multimasks": [low[f]], "multistep_pred_ious": [ious[f]]} for f in range(T)]
        fused(st, o2c, C, tg)["total_loss"].backward()

    def run_unfused(k):
        low, ious, tg = sets[k % nsets]
        unfused(low, ious, groups, tg, plain, S)["total_loss"].backward()

    def clear(k):
        low, ious, _ = sets[k % nsets]
        for v in low + ious:
            v.grad = None

    res = {}
    for nm, fn in (("fused", run_fused), ("unfused", run_unfused)):
        for k in range(3):
            fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(iters):
            fn(k)
            clear(k)
        e1.record()
        torch.cuda.synchronize()
        res[nm] = e0.elapsed_time(e1) / iters * 1e3
    # kernel-only times of the fused op (CUDA events around each C-ABI call)
    from sam2_video_training_b200 import ops
    ops.PROFILE = {}
    for k in range(iters):
        run_fused(k)
        clear(k)
    torch.cuda.synchronize()
    kern = {nm: sum(a.elapsed_time(b) for a, b, _ in v) / len(v) * 1e3 for nm, v in ops.PROFILE.items()}
    ops.PROFILE = None
    px = T * C * S * S
    fused_bytes = px * 2 + T * n_obj * s * s * 4 * 3          # targets fwd+bwd, low-res fwd read + bwd read + bwd write
    print(json.dumps({"case": name, "T": T, "C": C, "n_obj": n_obj, "S": S, "fused_us": round(res["fused"], 1), "fused_kernels_us": {k: round(v, 1) for k, v in kern.items()},
                      "unfused_us": round(res["unfused"], 1), "speedup": round(res["unfused"] / res["fused"], 2),
                      "category_Mpx": px / 1e6, "kernel_Gpx_per_s": round(px / sum(kern.values()) / 1e3, 1),
                      "fused_algorithmic_MB": round(fused_bytes / 1e6, 1),
                      "kernel_GBps": round(fused_bytes / sum(kern.values()) / 1e3, 1)}), flush=True)


if "--ncu" in sys.argv:      # short run for an ncu capture
    bench("1024 px, T=8, 4 categories x 2 objects", 8, 4, 2, 256, 2, 2)
    sys.exit(0)
bench("cfg2 clip (384 px, T=10, 7 categories x 2 objects)", 10, 7, 2, 96, 16, 32)
bench("cfg3 clip (512 px, T=8, 13 categories x 1 object)", 8, 13, 1, 128, 8, 32)
bench("1024 px, T=8, 4 categories x 2 objects", 8, 4, 2, 256, 6, 24)
bench("1024 px, T=8, 4 categories x 6 objects", 8, 4, 6, 256, 6, 24)
