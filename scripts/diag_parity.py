"""Diagnostic: per-tensor gradient cosine / rel-L2 of the B200 stack vs golden + oracle (GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import attention_oracle as ao, detgen
from sam2_video_training_b200.modeling.memory_attention import build_memory_attention

dev = torch.device("cuda:0")
def cos(a, b):
    a = a.detach().double().cpu().flatten(); b = b.detach().double().cpu().flatten()
    return float(a @ b / (a.norm() * b.norm()))
def rl2(a, b):
    a = a.detach().double().cpu().flatten(); b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm())

import itertools
FUSED = [True, False]
for tag, fused in itertools.product(["g4_b2_f2_p8", "g8_b3_f3_p12", "g12_b1_f1_p0"], FUSED):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"attn_{tag}.npz"))
    grid, batch, nf, nptr = int(g["grid"]), int(g["batch"]), int(g["n_frames"]), int(g["n_ptr"])
    model = build_memory_attention().to(dev).eval()
    model.use_fused_stack = fused
    params = detgen.det_params(detgen.param_shapes())
    with torch.no_grad():
        for n, p in model.named_parameters(): p.copy_(params[n])
    inp = {k: v.to(dev) for k, v in detgen.attention_inputs(grid, batch, nf, nptr).items()}
    lv = {k: inp[k].clone().requires_grad_(True) for k in ("curr", "curr_pos", "memory", "memory_pos")}
    out = model(lv["curr"], lv["memory"], lv["curr_pos"], lv["memory_pos"], nptr)
    out.backward(inp["grad_out"])
    print(tag, "fused" if fused else "composed", "out rel_l2", rl2(out, torch.from_numpy(g["out"])))
    for k in lv:
        print("   d_%s cos %.6f rel %.4f" % (k, cos(lv[k].grad, torch.from_numpy(g["d_" + k])), rl2(lv[k].grad, torch.from_numpy(g["d_" + k]))))
    for key in g.files:
        if key.startswith("dparam:"):
            pg = dict(model.named_parameters())[key[7:]].grad
            print("   %s cos %.6f" % (key, cos(pg, torch.from_numpy(g[key]))))

# cfg1-like random
params = ao.init_params(seed=0)
for (grid, b, nf, nptr), fused in itertools.product([(24, 1, 7, 28), (16, 2, 3, 12)], FUSED):
    n, m = grid * grid, nf * grid * grid + nptr
    gen = torch.Generator().manual_seed(7)
    curr = torch.randn(n, b, 256, generator=gen); curr_pos = torch.randn(n, b, 256, generator=gen) * 0.7
    memory = torch.randn(m, b, 64, generator=gen); memory_pos = torch.randn(m, b, 64, generator=gen) * 0.7
    gout = torch.randn(n, b, 256, generator=gen)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    lo = {k: v.clone().requires_grad_(True) for k, v in dict(curr=curr, curr_pos=curr_pos, memory=memory, memory_pos=memory_pos).items()}
    ref = ao.memory_attention(po, lo["curr"], lo["memory"], lo["curr_pos"], lo["memory_pos"], nptr)
    ref.backward(gout)
    model = build_memory_attention().to(dev).eval()
    model.use_fused_stack = fused
    with torch.no_grad():
        for nme, p in model.named_parameters(): p.copy_(params[nme])
    ld = {k: v.to(dev).clone().requires_grad_(True) for k, v in dict(curr=curr, curr_pos=curr_pos, memory=memory, memory_pos=memory_pos).items()}
    out = model(ld["curr"], ld["memory"], ld["curr_pos"], ld["memory_pos"], nptr)
    out.backward(gout.to(dev))
    print("random grid", grid, "fused" if fused else "composed", "out rel_l2", rl2(out, ref))
    ga = torch.cat([p.grad.flatten().cpu().double() for _, p in model.named_parameters()])
    gb = torch.cat([po[nme].grad.flatten().double() for nme, _ in model.named_parameters()])
    print("   global param-grad cos %.6f" % float(ga @ gb / (ga.norm() * gb.norm())))
    for k in ld: print("   d_%s cos %.6f" % (k, cos(ld[k].grad, lo[k].grad)))
    cs = sorted((cos(p.grad, po[nme].grad), nme) for nme, p in model.named_parameters())
    print("   worst param cos:", cs[:5])
