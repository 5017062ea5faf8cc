#!/usr/bin/env python
"""bench.py -- the hot path of SAM2 video fine-tuning on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic clips, per GPU:
  for every clip frame t = 1 .. T-1 (frame 0 does not run memory attention, sam2_base.py:680-684):
      MemoryAttention forward + backward for all objects of the batch, growing memory bank
      M_t = min(t,7) * (N + 4) keys, upstream gradient ~ N(0,1);
  MultiStepMultiMasksAndIous forward + backward over T frames x C objects x S^2 logits per clip;
  (N > 1 ranks) NCCL all-reduce of the 106 parameter gradients; fused AdamW step.
Workload at N = 1: BASELINE.json configs[1] -- 384 px (24x24 tokens), 10-frame clips, 7 memory frames
+ object pointers, 7 objects x 8 clips (B = 56), bf16 tensor-core math / fp32 accumulate.
Prints ONE JSON line (see the task contract).  `--impl reference` times the CPU oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "memory_attn_fwd_bwd_plus_mask_loss_clip_frames_per_sec"
UNIT = "clip-frames/s"

WORKLOADS = {
    # name: grid, frames T, objects per clip C, clips per GPU, image size S
    "cfg2_endovis18_384px_T10_7obj_x8clips": dict(grid=24, T=10, C=7, clips=8, S=384),
    "cfg1_384px_T10_1obj_x1clip": dict(grid=24, T=10, C=1, clips=1, S=384),
    "cfg3_cholec_512px_T8_13obj_x1clip": dict(grid=32, T=8, C=13, clips=1, S=512),
    "cfg4_1024px_T8_4obj_x1clip": dict(grid=64, T=8, C=4, clips=1, S=1024),
}
LOSS_W = {"loss_mask": 20, "loss_dice": 1, "loss_iou": 1, "loss_class": 0}
PER_CLIP_LOSS = False      # --per-clip-loss: one criterion call per clip (the reference trainer's pattern) instead of one for the step
# DRAM bytes (read + write) of the backward kernels of one cross-attention call at B=56, N=576, M=4060 from `ncu --set full`
# (profiles/r1_ncu_attn_v64_cfg2_cross.csv: dK 248.0 + dQ 177.0 MB); updated whenever a new capture is committed
TRAFFIC_CFG2_BYTES = 431.0e6
TRAFFIC_NOTE = ("per cross-attention backward call at B=56 N=576 M=4060 (dK + dQ kernels); algorithmic operand bytes q,k,mem,dO',dq,dk = "
                "0.30 GB on the raw-memory path (K is read once by each of the two kernels)")


def bank_sizes(t: int, n: int):
    """Frame t of a clip attends to the conditioning frame + the last 6 frames (num_maskmem = 7, sam2_base.py:551-596) and to
    the object pointers of up to 16 past frames, 4 tokens each (sam2_base.py:612-672): M = min(t, 7) N + 4 min(t, 16)."""
    nf, npt = min(t, 7), min(t, 16)
    return nf * n + 4 * npt, 4 * npt  # M, P


def algorithmic_flops(wl) -> float:
    """per object, per layer, fwd: self core 4N^2 d, cross core 4NMd, linears
    4(2Nd^2) + 2(2Nd^2) + 2(2M dm d), MLP 2(2N d ff); bwd = 2.5 x core + 2 x linears."""
    d, dm, ff, L = 256, 64, 2048, 4
    n = wl["grid"] ** 2
    b = wl["C"] * wl["clips"]
    tot = 0.0
    for t in range(1, wl["T"]):
        m, _ = bank_sizes(t, n)
        core = 4.0 * n * n * d + 4.0 * n * m * d
        lin = 4 * (2.0 * n * d * d) + 2 * (2.0 * n * d * d) + 2 * (2.0 * m * dm * d) + 2 * (2.0 * n * d * ff)
        tot += L * b * (3.5 * core + 3.0 * lin)
    return tot


def attention_core_flops(wl) -> float:
    d, L = 256, 4
    n = wl["grid"] ** 2
    b = wl["C"] * wl["clips"]
    return sum(L * b * 3.5 * (4.0 * n * n * d + 4.0 * n * bank_sizes(t, n)[0] * d) for t in range(1, wl["T"]))


def make_host_inputs(wl, seed, pin):
    """Per-step inputs as HOST tensors: per-frame current features, per-frame memory-encoder
    features (+ pointer tokens), positional encodings, mask logits / targets / IoU predictions."""
    g = torch.Generator().manual_seed(seed)
    n, b, T, C, S, clips = wl["grid"] ** 2, wl["C"] * wl["clips"], wl["T"], wl["C"], wl["S"], wl["clips"]

    def mk(*shape, scale=1.0, dtype=torch.float32):
        x = (torch.randn(*shape, generator=g) * scale).to(dtype)
        return x.pin_memory() if pin else x

    h = dict(
        curr=[mk(n, b, 256) for _ in range(1, T)],
        curr_pos=mk(n, b, 256, scale=0.7),
        # what the tracker keeps per processed frame (sam2_base.py:715-769, 795-811), in the reference's own layout:
        # memory-encoder features + their position encoding [B, 64, H, W] and the object pointer [B, 256]
        mem_feat=[mk(b, 64, wl["grid"], wl["grid"]) for _ in range(T - 1)],
        mem_pos=[mk(b, 64, wl["grid"], wl["grid"], scale=0.7) for _ in range(T - 1)],
        obj_ptr=[mk(b, 256) for _ in range(T - 1)],
        grad_out=[mk(n, b, 256) for _ in range(1, T)],
        logits=[mk(C, 1, S, S, scale=4.0) for _ in range(clips * T)],   # per-frame tensors, as the wrapper produces
        iou=[torch.rand(T, C, 1, generator=g) for _ in range(clips)],
    )
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    tg = []
    for c_i in range(clips):
        t = torch.zeros(T, C, S, S, dtype=torch.bool)
        for f in range(T):
            for ch in range(C):
                if C >= 4 and ch % 8 == 7:
                    continue  # 1 in 8 channels empty: exercises the valid filter
                cx = S * (0.25 + 0.5 * torch.rand((), generator=g))
                cy = S * (0.25 + 0.5 * torch.rand((), generator=g))
                ax = S * (0.08 + 0.2 * torch.rand((), generator=g))
                ay = S * (0.08 + 0.2 * torch.rand((), generator=g))
                t[f, ch] = ((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 < 1
        tg.append(t.pin_memory() if pin else t)
    h["targets"] = tg
    return h


def h2d_bytes(h) -> int:
    tot = 0
    for v in h.values():
        for x in (v if isinstance(v, list) else [v]):
            tot += x.numel() * x.element_size()
    return tot


def to_device(h, dev):
    out = {}
    for k, v in h.items():
        out[k] = [x.to(dev, non_blocking=True) for x in v] if isinstance(v, list) else v.to(dev, non_blocking=True)
    return out


class BankState:
    """The trainable tensors the bank assembly differentiates through (SAM2Base.maskmem_tpos_enc, obj_ptr_tpos_proj:
    sam2_base.py:138-141, 654-663) -- outside the MemoryAttention freeze map, their gradients are produced and dropped."""

    def __init__(self, dev):
        from sam2_video_training_b200 import memory_bank as mb
        self.mb = mb
        self.cfg = mb.BankConfig()
        g = torch.Generator().manual_seed(3)
        self.tpos = (torch.randn(7, 1, 1, 64, generator=g) * 0.02).to(dev).requires_grad_(True)
        self.proj = torch.nn.Linear(256, 64).to(dev)

    def output_dict(self, d, t):
        """Tracker state before frame t: frame 0 is the conditioning frame, frames 1 .. t-1 were tracked."""
        od = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        for f in range(t):
            od["cond_frame_outputs" if f == 0 else "non_cond_frame_outputs"][f] = {
                "maskmem_features": d["mem_feat"][f], "maskmem_pos_enc": [d["mem_pos"][f]], "obj_ptr": d["obj_ptr"][f]}
        return od

    def packed(self, d, wl, t):
        """memory_bank.assemble_memory_packed: the bank written once per frame in the kernels' layout (bf16, batch-first,
        tpos added on write, pointer tokens appended) -- two launches, inside the timed region."""
        return self.mb.assemble_memory_packed(self.cfg, t, self.output_dict(d, t), wl["T"], self.tpos, self.proj, training=True)

    def reference_layout(self, d, wl, t):
        """(memory, memory_pos, P) as SAM2Base._prepare_memory_conditioned_features builds them (for the reference legs)."""
        with torch.no_grad():
            m, p, n = self.mb.assemble_memory(self.cfg, t, self.output_dict(d, t), wl["T"], self.tpos, self.proj, training=True)
        return m, p.detach().requires_grad_(True), n

    def drop_grads(self):
        self.tpos.grad = None
        self.proj.zero_grad(set_to_none=True)


def to_device_streamed(h, dev, wl, stream):
    """H2D of one step's inputs on `stream` in the order the step consumes them, with one event per attention frame
    and one for the loss inputs, so that the first frame starts as soon as ITS inputs have landed."""
    out = {k: ([None] * len(v) if isinstance(v, list) else None) for k, v in h.items()}
    events = []
    with torch.cuda.stream(stream):
        out["curr_pos"] = h["curr_pos"].to(dev, non_blocking=True)
        for t in range(1, wl["T"]):
            out["curr"][t - 1] = h["curr"][t - 1].to(dev, non_blocking=True)
            out["grad_out"][t - 1] = h["grad_out"][t - 1].to(dev, non_blocking=True)
            for k in ("mem_feat", "mem_pos", "obj_ptr"):        # outputs of frame t - 1: first read by frame t
                out[k][t - 1] = h[k][t - 1].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            events.append(ev)
        for k in ("logits", "iou", "targets"):
            out[k] = [x.to(dev, non_blocking=True) for x in h[k]]
        ev = torch.cuda.Event()
        ev.record(stream)
        events.append(ev)
    return out, events


def run_step(model, crit, opt, d, bank, wl, world, fwd=None, arrivals=None):
    """`bank`: BankState -- the memory bank of every frame is assembled HERE, inside the step, from the per-frame tracker
    outputs (packed layout).  `arrivals` (end-to-end leg): events of the copy stream, one per attention frame (inputs of
    frame t and the memory frames it reads have landed) + one for the loss inputs."""
    T, C = wl["T"], wl["C"]
    fwd = fwd or model
    for t in range(1, T):
        if arrivals is not None:
            torch.cuda.current_stream().wait_event(arrivals[t - 1])
        out = fwd(d["curr"][t - 1], bank.packed(d, wl, t), d["curr_pos"])
        out.backward(d["grad_out"][t - 1])
    bank.drop_grads()
    # the parameter gradients are complete: the all-reduce starts here and overlaps the (parameter-free) mask loss below
    pending = None
    if world > 1:
        from sam2_video_training_b200 import ddp
        pending = ddp.allreduce_gradients_async(model, world)
    total = None
    if arrivals is not None:
        torch.cuda.current_stream().wait_event(arrivals[-1])
    # the criterion over ALL clips of the step in one launch (forward_clips: one target pointer per frame); the per-clip form is
    # what the reference trainer does (trainer.py:268) and what --per-clip-loss times
    clips, leaves = [], []
    for ci in range(wl["clips"]):
        xs = [d["logits"][ci * T + f].requires_grad_(True) for f in range(T)]
        # per-frame IoU-head outputs are separate leaves, as the tracker produces them (one [C, 1] tensor per frame)
        ips = [v.detach().requires_grad_(True) for v in d["iou"][ci].unbind(0)]
        outs = [{"multistep_pred_multimasks_high_res": [xs[f]], "multistep_pred_ious": [ips[f]],
                 "multistep_object_score_logits": [None]} for f in range(T)]
        clips.append((outs, d["targets"][ci]))
        leaves += xs
    if PER_CLIP_LOSS:
        for outs, tg in clips:
            losses = crit(outs, tg)
            losses["total_loss"].backward()
            total = losses["total_loss"].detach() if total is None else total + losses["total_loss"].detach()
    else:
        losses = crit.forward_clips(clips)
        losses["total_loss"].backward()
        total = losses["total_loss"].detach()
    for x in leaves:
        x.grad = None
    if pending is not None:
        pending.wait()
    opt.step()
    model._sam2b200_grad_bucket.zero()   # one memset for all 106 gradients
    return total


class ClockSampler:
    def __init__(self, dev_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev_index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        res = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return res
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons, mx = [], set(), None
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            # under load = top half of the samples
            sm.sort()
            res["sm_mhz"] = sm[len(sm) // 2 + len(sm) // 4] if len(sm) > 3 else sm[-1]
        res["sm_max_mhz"] = mx
        res["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return res


def loss_sweep_roofline(lib, dev, hbm_peak, C=32, S=1024, nsets=3, iters=12):
    """Fused mask loss fwd + bwd at the BASELINE.json sweep maximum (32 objects x 1024^2 logits) through the C ABI:
    `nsets` independent input sets (3 x 168 MB of logits > 126 MB L2) used round-robin, CUDA events around the loop."""
    from sam2_video_training_b200 import _lib
    hw = S * S
    g = torch.Generator(device="cuda").manual_seed(7)
    st = torch.cuda.current_stream().cuda_stream
    sets = []
    yy, xx = torch.meshgrid(torch.arange(S, device=dev), torch.arange(S, device=dev), indexing="ij")
    for i in range(nsets):
        x = torch.randn(C, hw, device=dev, generator=g) * 4
        tg = torch.zeros(1, C, S, S, dtype=torch.uint8, device=dev)
        for c in range(C):
            if c % 8 != 7:
                cx, cy, ax, ay = [float(v) for v in torch.rand(4, generator=torch.Generator().manual_seed(97 * i + c))]
                tg[0, c] = (((xx - S * (.25 + .5 * cx)) / (S * (.08 + .2 * ax))) ** 2 + ((yy - S * (.25 + .5 * cy)) / (S * (.08 + .2 * ay))) ** 2 < 1)
        dl = torch.empty(C, hw, device=dev)
        sets.append(dict(x=x, tg=tg, dl=dl, lp=_lib.ptr_array([x.data_ptr()]), dp=_lib.ptr_array([dl.data_ptr()]),
                         iou=torch.rand(1, C, device=dev, generator=g), sums=torch.empty(1, C, 6, device=dev),
                         nv=torch.empty(1, dtype=torch.int32, device=dev), losses=torch.zeros(4, device=dev),
                         diou=torch.empty(1, C, device=dev),
                         ws=torch.zeros(max(lib.sam2b200_mask_loss_workspace_bytes(1, C, hw), 4) // 4, device=dev)))
    gl = torch.tensor([20.0, 1.0, 1.0, 0.0], device=dev)

    def fwd(s_):
        _lib.check(lib.sam2b200_mask_loss_fwd(s_["lp"], s_["tg"].data_ptr(), s_["iou"].data_ptr(), None, s_["ws"].data_ptr(),
                                              s_["sums"].data_ptr(), s_["nv"].data_ptr(), s_["losses"].data_ptr(), 1, C, hw, 0x100,
                                              0.25, 2.0, 1.0, 1, 1, st), "mask_loss_fwd")

    def bwd(s_):
        _lib.check(lib.sam2b200_mask_loss_bwd(s_["lp"], s_["dp"], s_["tg"].data_ptr(), s_["iou"].data_ptr(), None, s_["sums"].data_ptr(),
                                              s_["nv"].data_ptr(), gl.data_ptr(), s_["diou"].data_ptr(), 1, C, hw, 0, 0.25, 2.0, 1.0,
                                              1, 1, st), "mask_loss_bwd")

    def timeit(fn):
        for s_ in sets:
            fn(s_)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(sets[i % nsets])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    tf, tb = timeit(fwd), timeit(bwd)
    px = C * hw
    ach = 14.0 * px / ((tf + tb) * 1e-3) / 1e9
    return {"bound": "hbm", "workload": "%d objects x %d^2 logits, fwd + bwd, %d rotating input sets (inputs larger than L2)" % (C, S, nsets),
            "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "bytes_per_px": "5 fwd + 9 bwd",
            "fwd_us": tf * 1e3, "bwd_us": tb * 1e3, "fwd_gbs": 5.0 * px / (tf * 1e-3) / 1e9, "bwd_gbs": 9.0 * px / (tb * 1e-3) / 1e9}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ----------------------------------------------------------------------------- the reference itself (baseline/_ref)
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")


def load_reference():
    """The UNMODIFIED reference modules of the hot path from the git-ignored baseline/_ref/ (placed there by
    scripts/install_reference.py; the contract's pip install fails -- the reference has no setup.py / pyproject).
    They import the un-installed pip package `sam2`, so the vendored files are registered under those names
    (sam2_video/model/modeling/memory_attention.py:12-14).  Returns a namespace or None when baseline/_ref is absent."""
    import importlib.util
    import types
    mdl = os.path.join(REF_ROOT, "sam2_video", "model", "modeling")
    if not os.path.isfile(os.path.join(mdl, "memory_attention.py")):
        return None
    if "_bench_ref_ns" in sys.modules:
        return sys.modules["_bench_ref_ns"]

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    for pkg in ("sam2", "sam2.modeling", "sam2.modeling.sam", "sam2.utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    if "sam2.utils.misc" not in sys.modules:
        misc = types.ModuleType("sam2.utils.misc")
        misc.mask_to_box = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("outside the hot path"))
        sys.modules["sam2.utils.misc"] = misc
    load("sam2.modeling.position_encoding", os.path.join(mdl, "position_encoding.py"))
    load("sam2.modeling.sam2_utils", os.path.join(mdl, "sam2_utils.py"))
    tr = load("sam2.modeling.sam.transformer", os.path.join(mdl, "sam", "transformer.py"))
    ma = load("sam2.modeling.memory_attention", os.path.join(mdl, "memory_attention.py"))
    ls = load("_bench_ref_losses", os.path.join(REF_ROOT, "sam2_video", "model", "losses.py"))
    try:
        from loguru import logger
        logger.disable("_bench_ref_losses")
    except Exception:
        pass
    ns = types.SimpleNamespace(transformer=tr, memory_attention=ma, losses=ls)
    sys.modules["_bench_ref_ns"] = ns
    return ns


def build_reference_stack(ns, dropout=0.0):
    """configs/sam2/sam2.1_hiera_t.yaml:29-60 through the reference's own constructors."""
    sa = ns.transformer.RoPEAttention(rope_theta=10000.0, feat_sizes=[64, 64], embedding_dim=256, num_heads=1,
                                      downsample_rate=1, dropout=dropout)
    ca = ns.transformer.RoPEAttention(rope_theta=10000.0, feat_sizes=[64, 64], rope_k_repeat=True, embedding_dim=256,
                                      num_heads=1, downsample_rate=1, dropout=dropout, kv_in_dim=64)
    layer = ns.memory_attention.MemoryAttentionLayer(activation="relu", dim_feedforward=2048, dropout=dropout, pos_enc_at_attn=False,
                                                     self_attention=sa, d_model=256, pos_enc_at_cross_attn_keys=True,
                                                     pos_enc_at_cross_attn_queries=False, cross_attention=ca)
    return ns.memory_attention.MemoryAttention(d_model=256, pos_enc_at_input=True, layer=layer, num_layers=4)


def reference_clip_sample(ns, wl, threads, objects=None):
    """ONE clip of the workload (all `objects` = C objects batched along B as the reference does, dataset.py:358) through
    the reference's own MemoryAttention (T-1 frames fwd+bwd, growing bank) and MultiStepMultiMasksAndIous (T frames
    fwd+bwd) on the host cores, fp32.  Returns seconds."""
    torch.set_num_threads(threads)
    n, T, S = wl["grid"] ** 2, wl["T"], wl["S"]
    C = objects or wl["C"]
    g = torch.Generator().manual_seed(11)
    if "_ref_cpu_model" not in reference_clip_sample.__dict__:
        torch.manual_seed(0)
        reference_clip_sample._ref_cpu_model = build_reference_stack(ns, 0.0).train()
    model = reference_clip_sample._ref_cpu_model
    crit = ns.losses.MultiStepMultiMasksAndIous(weight_dict=dict(LOSS_W), supervise_all_iou=True, iou_use_l1_loss=True,
                                                pred_obj_scores=False, focal_gamma_obj_score=0.0, focal_alpha_obj_score=-1.0)
    curr_pos = torch.randn(n, C, 256, generator=g) * 0.7
    t0 = time.perf_counter()
    model.zero_grad(set_to_none=True)
    for t in range(1, T):
        m, p = bank_sizes(t, n)
        curr = torch.randn(n, C, 256, generator=g)
        mem = torch.randn(m, C, 64, generator=g)
        pos = (torch.randn(m, C, 64, generator=g) * 0.7).requires_grad_(True)
        out = model(curr=[curr], curr_pos=[curr_pos], memory=mem, memory_pos=pos, num_obj_ptr_tokens=p)
        out.backward(torch.randn(n, C, 256, generator=g))
    logits = (torch.randn(T, C, 1, S, S, generator=g) * 4).requires_grad_(True)
    targets = torch.rand(T, C, S, S, generator=g) > 0.8
    iou = torch.rand(T, C, 1, generator=g).requires_grad_(True)
    outs = [{"multistep_pred_multimasks_high_res": [logits[f]], "multistep_pred_ious": [iou[f]],
             "multistep_object_score_logits": [torch.zeros(C, 1)]} for f in range(T)]
    crit(outs, targets)["total_loss"].backward()
    return time.perf_counter() - t0


def cpu_oracle_sample(wl, threads):
    """Fallback when baseline/_ref is absent: one object-clip through the CPU oracle port (oracle/), fp32.  Seconds."""
    from oracle import attention_oracle as ao
    from oracle import losses_oracle as lo
    torch.set_num_threads(threads)
    n, T, S = wl["grid"] ** 2, wl["T"], wl["S"]
    g = torch.Generator().manual_seed(11)
    params = {k: v.clone().requires_grad_(True) for k, v in ao.init_params(seed=0).items()}
    curr_pos = torch.randn(n, 1, 256, generator=g) * 0.7
    t0 = time.perf_counter()
    for t in range(1, T):
        m, p = bank_sizes(t, n)
        curr = torch.randn(n, 1, 256, generator=g)
        mem = torch.randn(m, 1, 64, generator=g)
        pos = torch.randn(m, 1, 64, generator=g) * 0.7
        out = ao.memory_attention(params, curr, mem, curr_pos, pos, p)
        out.backward(torch.randn(n, 1, 256, generator=g))
    logits = (torch.randn(T, 1, 1, S, S, generator=g) * 4).requires_grad_(True)
    targets = torch.rand(T, 1, S, S, generator=g) > 0.8
    iou = torch.rand(T, 1, 1, generator=g).requires_grad_(True)
    l = lo.multistep_loss([logits[f] for f in range(T)], targets, [iou[f] for f in range(T)], dict(LOSS_W), iou_use_l1_loss=True)
    l["total_loss"].backward()
    return time.perf_counter() - t0


def cpu_reference_leg(wl, budget_s, min_samples=2, max_samples=64):
    """The reference's CPU path on all host cores for ~budget_s seconds.  Returns (clip-frames/s, dict for the JSON line)."""
    threads = os.cpu_count() or 1
    ns = load_reference()
    if ns is not None:
        reference_clip_sample(ns, WORKLOADS["cfg1_384px_T10_1obj_x1clip"], threads)     # warm-up (thread pool, allocator)
        ts, t_start = [], time.perf_counter()
        while len(ts) < min_samples or (time.perf_counter() - t_start < budget_s and len(ts) < max_samples):
            ts.append(reference_clip_sample(ns, wl, threads))
        tb = sum(ts) / len(ts)
        value = wl["T"] / tb            # one clip = T clip-frames
        return value, tb, {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                           "sample": "the UNMODIFIED reference modules (baseline/_ref: MemoryAttention + MultiStepMultiMasksAndIous, torch CPU "
                                     "fp32, dropout 0): %d clip(s) of the %d per step, each = %d objects batched x (%d attention frames fwd+bwd + "
                                     "loss on %d x %d x %d^2); mean of %d after 1 warm-up (%.2f s each)" % (
                                         len(ts), wl["clips"], wl["C"], wl["T"] - 1, wl["T"], wl["C"], wl["S"], len(ts), tb)}
    cpu_oracle_sample(WORKLOADS["cfg1_384px_T10_1obj_x1clip"], threads)
    ts, t_start = [], time.perf_counter()
    while len(ts) < min_samples or (time.perf_counter() - t_start < budget_s and len(ts) < max_samples):
        ts.append(cpu_oracle_sample(wl, threads))
    tb = sum(ts) / len(ts)
    value = (wl["T"] / wl["C"]) / tb    # one object-clip = T / C clip-frames
    return value, tb, {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": "baseline/_ref absent -> oracle/ port, torch CPU fp32: %d object-clips (of %d per step), scaled by 1/C "
                                 "to clip-frames (%.2f s each)" % (len(ts), wl["C"] * wl["clips"], tb)}


def same_box_torch_gpu(wl, dev, d, banks, reps=2):
    """The SAME step (attention fwd+bwd for every frame of every object + the loss, no optimizer) through the unmodified
    reference modules moved to this GPU: torch SDPA + eager loss, fp32 and bf16-autocast.  This is the bar SURVEY.md
    section 2.2 names (the reference has no Blackwell kernel of its own).  None when baseline/_ref is absent."""
    ns = load_reference()
    if ns is None:
        return None
    torch.manual_seed(0)
    model = build_reference_stack(ns, 0.0).to(dev).train()
    crit = ns.losses.MultiStepMultiMasksAndIous(weight_dict=dict(LOSS_W), supervise_all_iou=True, iou_use_l1_loss=True,
                                                pred_obj_scores=False, focal_gamma_obj_score=0.0, focal_alpha_obj_score=-1.0)
    T, C = wl["T"], wl["C"]

    ref_banks = [banks.reference_layout(d, wl, t) for t in range(1, T)]     # built once, outside the timed region

    def step(autocast):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            for t in range(1, T):
                mem, pos, p = ref_banks[t - 1]
                out = model(curr=[d["curr"][t - 1]], curr_pos=[d["curr_pos"]], memory=mem, memory_pos=pos, num_obj_ptr_tokens=p)
                out.backward(d["grad_out"][t - 1].to(out.dtype))
                pos.grad = None
            for ci in range(wl["clips"]):
                xs = [d["logits"][ci * T + f].requires_grad_(True) for f in range(T)]
                ips = [v.detach().requires_grad_(True) for v in d["iou"][ci].unbind(0)]
                outs = [{"multistep_pred_multimasks_high_res": [xs[f]], "multistep_pred_ious": [ips[f]],
                         "multistep_object_score_logits": [torch.zeros(C, 1, device=dev)]} for f in range(T)]
                crit(outs, d["targets"][ci])["total_loss"].backward()
                for x in xs:
                    x.grad = None

    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, ac in (("bf16_autocast_ms", True), ("fp32_ms", False)):
        step(ac)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            step(ac)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps
    res["what"] = ("unmodified reference MemoryAttention (F.scaled_dot_product_attention) + MultiStepMultiMasksAndIous from baseline/_ref on "
                   "this GPU, same inputs and step as `value` minus the optimizer; torch %s" % torch.__version__)
    del model
    torch.cuda.empty_cache()
    return res


def main_reference(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ns = load_reference()
    sample = (lambda: reference_clip_sample(ns, wl, threads)) if ns is not None else (lambda: cpu_oracle_sample(wl, threads))
    per_sample_frames = wl["T"] if ns is not None else wl["T"] / wl["C"]
    for _ in range(min(max(args.warmup, 0), 1)):      # one warm-up (thread pool, allocator): there are no clocks to ramp on the CPU,
        sample()                                       # and a sample is seconds long
    ts = [sample() for _ in range(max(args.steps, 1))]
    tmean = sum(ts) / len(ts)
    value = per_sample_frames / tmean
    kind = "reference" if ns is not None else "port"
    what = ("the UNMODIFIED reference modules from baseline/_ref (torch CPU fp32, dropout 0): 1 clip per step = %d objects batched x (%d "
            "attention frames fwd+bwd + loss on %d frames); the workload has %d such clips per step" % (wl["C"], wl["T"] - 1, wl["T"], wl["clips"])
            if ns is not None else "baseline/_ref absent -> oracle/ port: 1 object-clip per step, scaled by 1/C to clip-frames")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tmean * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "sample_per_step": what},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def measure_workload(wl_name, wl, model, crit, opt, lib, dev, world, rank, steps, warmup, use_graphs, barrier, pk, want_clocks):
    """Device-resident throughput of one workload + the per-kernel-family pass.  Returns a dict (and keeps the device
    inputs in it for the callers that go on to the end-to-end / same-box legs)."""
    import torch.distributed as dist
    from sam2_video_training_b200 import fused_stack, ops
    from sam2_video_training_b200.graphs import GraphedMemoryAttention
    run_model = GraphedMemoryAttention(model) if use_graphs else model
    host = make_host_inputs(wl, 1234 + rank, pin=True)
    d = to_device(host, dev)
    banks = BankState(dev)
    torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        run_step(model, crit, opt, d, banks, wl, world, fwd=run_model)
    crit.raise_if_invalid()
    barrier()
    clocks = ClockSampler(dev.index) if want_clocks else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t_host0 = time.perf_counter()
    for _ in range(steps):
        run_step(model, crit, opt, d, banks, wl, world, fwd=run_model)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / steps     # host time to ENQUEUE a step (no synchronisation inside)
    ev1.record()
    barrier()
    crit.raise_if_invalid()          # the reference's "No valid masks" contract, checked once outside the timed region
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / steps
    clk = clocks.stop() if clocks is not None else None

    # same steps, every kernel launched from the host on ONE stream, CUDA events around each kernel family (events cannot
    # be timed inside a replayed graph, and with the side stream an event pair around one kernel would time its neighbours)
    fused_stack.NO_SIDE_STREAM = True
    run_step(model, crit, opt, d, banks, wl, world)
    barrier()
    launches0 = lib.sam2b200_launch_count()
    ops.PROFILE = {}
    ev0.record()
    for _ in range(steps):
        run_step(model, crit, opt, d, banks, wl, world)
    ev1.record()
    barrier()
    prof, ops.PROFILE = ops.PROFILE, None
    fused_stack.NO_SIDE_STREAM = bool(os.environ.get("SAM2B200_NO_SIDE_STREAM"))
    eager_ms = ev0.elapsed_time(ev1) / steps
    launches = lib.sam2b200_launch_count() - launches0      # a replayed graph launches the same kernels
    crit.raise_if_invalid()
    fam = {}
    for name, evs in prof.items():
        fam[name] = dict(ms=sum(a.elapsed_time(b) for a, b, _ in evs), work=sum(w for _, _, w in evs), launches=len(evs))
    attn_ms = sum(fam[k]["ms"] for k in ("attn_fwd", "attn_bwd") if k in fam)
    attn_fl = sum(fam[k]["work"] for k in ("attn_fwd", "attn_bwd") if k in fam)
    loss_ms = sum(fam[k]["ms"] for k in ("mask_loss_fwd", "mask_loss_bwd") if k in fam)
    loss_by = sum(fam[k]["work"] for k in ("mask_loss_fwd", "mask_loss_bwd") if k in fam)
    dom = max(("attn_fwd", "attn_bwd"), key=lambda k: fam.get(k, {"ms": 0})["ms"])
    ach = fam[dom]["work"] / (fam[dom]["ms"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "sam2b200_" + dom + " (tcgen05 kernels, all launches of the timed region)",
                "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                "peak_source": pk["src"] + " sustained cuBLAS bf16",
                "timing_note": "kernel time from CUDA events on the launching stream in a second, single-stream, host-launched pass of "
                               "the same steps (events cannot sit inside the graph-replayed timed region)",
                "frac_of_spec_2250": ach / 2250.0,
                "attention_fwd_bwd_tflops": attn_fl / (attn_ms * 1e-3) / 1e12 if attn_ms else None,
                "attention_fwd_tflops": (fam["attn_fwd"]["work"] / (fam["attn_fwd"]["ms"] * 1e-3) / 1e12) if "attn_fwd" in fam else None,
                "attention_share_of_step": attn_ms / (eager_ms * steps) if eager_ms else None,
                "mask_loss": {"bound": "hbm", "achieved": loss_by / (loss_ms * 1e-3) / 1e9 if loss_ms else None, "peak": pk["hbm"],
                              "unit": "GB/s", "frac": (loss_by / (loss_ms * 1e-3) / 1e9 / pk["hbm"]) if loss_ms else None,
                              "bytes_per_px": "5 fwd + 9 bwd"},
                "whole_step_tflops": algorithmic_flops(wl) / (ms_per_step * 1e-3) / 1e12}
    return dict(ms_per_step=ms_per_step, clocks=clk, eager_ms=eager_ms, host_enqueue_ms=host_enqueue_ms, launches=int(launches), fam=fam, roofline=roofline,
                host=host, d=d, banks=banks, run_model=run_model)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_endovis18_384px_T10_7obj_x8clips", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_workloads / same_box_torch_gpu / the dropout leg")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host (no CUDA-graph replay)")
    ap.add_argument("--per-clip-loss", action="store_true", help="one criterion call per clip instead of one launch for all clips of the step")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="train-mode dropout of the stack (the reference ships 0.1); default 0 = the parity configuration "
                         "SURVEY.md section 8d prescribes for the headline number")
    args = ap.parse_args()
    global PER_CLIP_LOSS
    PER_CLIP_LOSS = bool(args.per_clip_loss)
    wl_name, wl = args.workload, WORKLOADS[args.workload]
    if args.impl == "reference":
        return main_reference(args, wl_name, wl)

    import torch.distributed as dist
    from sam2_video_training_b200 import _lib, ddp
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    _lib.check(lib.sam2b200_check_device(local_rank), "sam2b200_check_device")

    torch.manual_seed(0)
    model = build_memory_attention(dropout=args.dropout).to(dev).train()   # default 0: throughput with parity numerics
    # the reference's "No valid masks" contract stays ON: the per-frame valid-channel minimum is folded on the device and
    # checked at the step's 4-byte read-back (e2e leg) / once per timed region (device-resident leg) -- no host sync per call
    crit = MultiStepMultiMasksAndIous(dict(LOSS_W), supervise_all_iou=True, iou_use_l1_loss=True, check_valid="deferred")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True)
    ddp.attach_grad_bucket(model)   # all 106 gradients are views of one flat fp32 buffer
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    r = measure_workload(wl_name, wl, model, crit, opt, lib, dev, world, rank, args.steps, args.warmup, not args.no_graphs,
                         barrier, pk, want_clocks=(rank == 0))
    ms_per_step, clk, host, d, banks, run_model = r["ms_per_step"], r["clocks"], r["host"], r["d"], r["banks"], r["run_model"]
    roofline, fam = r["roofline"], r["fam"]
    r_launches, eager_ms_headline, host_enqueue_ms = r["launches"], r["eager_ms"], r["host_enqueue_ms"]
    frames_per_step = wl["T"] * wl["clips"]
    value = world * frames_per_step / (ms_per_step * 1e-3)
    # dram__bytes_read.sum + dram__bytes_write.sum of the backward kernels of ONE cross-attention call at the largest shape of
    # cfg2 (B=56, N=576, M=4060), ncu --set full: see profiles/ (r2 file when present, else the round-1 capture)
    roofline["traffic"] = (TRAFFIC_CFG2_BYTES if wl_name.startswith("cfg2") else None)
    roofline["traffic_note"] = TRAFFIC_NOTE
    if rank == 0:
        roofline["mask_loss_sweep_max"] = loss_sweep_roofline(lib, dev, pk["hbm"])
        roofline["mask_loss"]["note"] = ("in-step: %d per-clip calls of %d x %d x %d^2" % (wl["clips"], wl["T"], wl["C"], wl["S"]) if PER_CLIP_LOSS else
                                         "in-step: ONE launch for the %d clips of the step (%d frames x %d x %d^2)" % (wl["clips"], wl["clips"] * wl["T"], wl["C"], wl["S"]))

    # ---------------- end to end: host buffers, H2D inside the timed region, loss read back ----------------
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()

        def start_copy():   # H2D of one step's inputs from pinned host memory, on the copy stream, in consumption order
            return to_device_streamed(host, dev, wl, copy_stream)

        def e2e_loop(k):
            # double-buffered: the copy of step i+1 overlaps the compute of step i, and within a step frame t only waits
            # for its own inputs; every copy is inside the timed region
            nxt = start_copy()
            for i in range(k):
                dd, evs = nxt
                tot = run_step(model, crit, opt, dd, banks, wl, world, fwd=run_model, arrivals=evs)
                if i + 1 < k:       # enqueue the next step's copies AFTER this step's kernels: the compute stream never
                    nxt = start_copy()   # waits for the host to issue ~130 cudaMemcpyAsync calls
                # D2H of the step's result: [loss, min valid channels] in ONE 8-byte read (also keeps `dd` alive until the
                # step is done); the validity contract of the reference is enforced from it
                both = torch.stack([tot.float(), crit.deferred_state().float()]).cpu()
                crit.raise_if_invalid(value=int(both[1]))

        e2e_loop(2)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        e2e_loop(args.steps)
        ev1.record()
        barrier()
        t2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item()) / args.steps
        e2e = {"value": world * frames_per_step / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(host),
               "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms}

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        # ---- the bar SURVEY.md 2.2 names: the reference modules on this same GPU (torch SDPA), same step
        try:
            sb = same_box_torch_gpu(wl, dev, d, banks)
        except Exception as e:      # the baseline leg must never take the measurement down
            sb = {"error": repr(e)[:300]}
        if sb is not None:
            if "bf16_autocast_ms" in sb:
                sb["speedup_vs_bf16_autocast"] = sb["bf16_autocast_ms"] / ms_per_step
                sb["speedup_vs_fp32"] = sb["fp32_ms"] / ms_per_step
            extras["same_box_torch_gpu"] = sb
        # ---- the shipped training configuration has dropout 0.1 in the stack: same step with it switched on
        if args.dropout == 0.0:
            try:
                for mod in model.modules():
                    if isinstance(mod, torch.nn.Dropout):
                        mod.p = 0.1
                for layer in model.layers:
                    layer.dropout_value = 0.1
                    layer.self_attn.dropout_p = 0.1
                    layer.cross_attn_image.dropout_p = 0.1
                rd = measure_workload(wl_name, wl, model, crit, opt, lib, dev, world, rank, max(2, args.steps // 2), 3, not args.no_graphs,
                                      barrier, pk, want_clocks=False)
                extras["dropout_0.1"] = {"ms_per_step": rd["ms_per_step"], "value": frames_per_step / (rd["ms_per_step"] * 1e-3),
                                         "overhead_vs_dropout_0": rd["ms_per_step"] / ms_per_step - 1.0}
                del rd
            finally:
                for mod in model.modules():
                    if isinstance(mod, torch.nn.Dropout):
                        mod.p = 0.0
                for layer in model.layers:
                    layer.dropout_value = 0.0
                    layer.self_attn.dropout_p = 0.0
                    layer.cross_attn_image.dropout_p = 0.0
    del d, banks, run_model, r
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_extras:
        # ---- the other single-GPU shapes of BASELINE.json (cfg3: 512 px x 13 objects; cfg4: 1024 px x 4 objects)
        other = {}
        for name in ("cfg3_cholec_512px_T8_13obj_x1clip", "cfg4_1024px_T8_4obj_x1clip"):
            if name == wl_name:
                continue
            try:
                ro = measure_workload(name, WORKLOADS[name], model, crit, opt, lib, dev, world, rank, 3, 3, not args.no_graphs, barrier, pk,
                                      want_clocks=False)
                w2 = WORKLOADS[name]
                other[name] = {"ms_per_step": ro["ms_per_step"], "value": w2["T"] * w2["clips"] / (ro["ms_per_step"] * 1e-3), "unit": UNIT,
                               "algorithmic_tflop_per_step": algorithmic_flops(w2) / 1e12,
                               "roofline": {k: ro["roofline"][k] for k in ("kernel", "achieved", "frac", "peak", "frac_of_spec_2250",
                                                                          "attention_fwd_bwd_tflops", "attention_fwd_tflops", "whole_step_tflops")},
                               "kernel_families_ms_per_step": {k: v["ms"] / 3 for k, v in ro["fam"].items()}}
                del ro
            except Exception as e:
                other[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        extras["other_workloads"] = other

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu_baseline = cpu_reference_leg(wl, budget_s=12.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": wl_name, "tokens": wl["grid"] ** 2, "frames_per_clip": wl["T"], "objects_per_clip": wl["C"],
                       "clips_per_gpu": wl["clips"], "mask_px": wl["S"], "memory_bank": "frame t: min(t,7) memory frames x N tokens + 4 x min(t,16) pointer tokens, assembled per frame inside the timed region (memory_bank.assemble_memory_packed)",
                       "l2": "inputs_larger_than_L2 (%.0f MB per step)" % (h2d_bytes(host) / 1e6),
                       "launch": "memory-attention fwd/bwd replayed as CUDA graphs (one pair per memory-bank shape)" if not args.no_graphs else "host-launched",
                       "parallelism": "dp%d (clips sharded, NCCL grad all-reduce)" % world,
                       "dropout": args.dropout, "loss_validity_check": "on (deferred: device flag read with the loss)",
                       "object_frames_per_step": wl["T"] * wl["C"] * wl["clips"],
                       "algorithmic_tflop_per_step": algorithmic_flops(wl) / 1e12},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(r_launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernel_families_ms_per_step": {k: v["ms"] / args.steps for k, v in fam.items()},
            "ms_per_step_host_launched": eager_ms_headline, "host_enqueue_ms_per_step": host_enqueue_ms, "cuda_graphs": not args.no_graphs,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
