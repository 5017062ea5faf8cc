"""One call of each round-2 kernel at the cfg2 shapes (B = 56, N = 576, M = 4060), for `ncu --set full -k regex:...` captures."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ops, fused_stack as fs
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0")
BF16 = torch.bfloat16
b, n, m = 56, 576, 4060
g = torch.Generator(device="cuda").manual_seed(1)
q = (torch.randn(b, n, 256, device=dev, generator=g) * 1.5).to(BF16)
k = (torch.randn(b, m, 256, device=dev, generator=g) * 1.5).to(BF16)
ks = (torch.randn(b, n, 256, device=dev, generator=g) * 1.5).to(BF16)
vs = torch.randn(b, n, 256, device=dev, generator=g).to(BF16)
mem = torch.randn(b, m, 64, device=dev, generator=g).to(BF16)
do = torch.randn(b, n, 256, device=dev, generator=g).to(BF16)
do64 = torch.randn(b, n, 64, device=dev, generator=g).to(BF16)
wo = (torch.randn(256, 256, device=dev, generator=g) / 16).to(BF16)
weff = (torch.randn(256, 64, device=dev, generator=g) / 8).to(BF16)
bias = torch.zeros(256, device=dev)
table = compute_axial_cis(dim=256, end_x=24, end_y=24).to(dev)
nr = (m // n) * n
for it in range(2):
    # forward kernels with the fused output projection
    o, o32, lse, sa = ops.attn_fwd_proj(q, ks, vs, wo, bias, 1 / 16.0)
    o64, o64_32, lse64, _, ca = ops.attn_fwd_v64_proj(q, k, mem, weff, bias, None, 1 / 16.0)
    delta = (do64.float() * o64_32).sum(-1)
    db = tuple(torch.zeros(256, device=dev) for _ in range(3))
    # backward kernels, real-call mode
    ops.attn_bwd_v64(q, k, mem, do64, lse64, delta, 1 / 16.0, table=table, n_rope_k=nr, grad_dtype=BF16, dbias=db[:2])
    ops.attn_bwd(q, ks, vs, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=n, grad_dtype=BF16, dbias=db)
    # LayerNorm + projection heads
    x = torch.randn(b * n, 256, device=dev, generator=g); res = torch.randn(b * n, 256, device=dev, generator=g).to(BF16)
    gm, bt = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    wqkv = (torch.randn(768, 256, device=dev, generator=g) / 16).to(BF16); w1 = (torch.randn(2048, 256, device=dev, generator=g) / 16).to(BF16)
    fs.ln_proj(x, res, gm, bt, wqkv, torch.zeros(768, device=dev, dtype=BF16), 3, table=table, rope_outs=2, rows_per_item=n, n_rope_rows=n)
    fs.ln_proj(x, res, gm, bt, w1, torch.zeros(2048, device=dev, dtype=BF16), 1, out_width=2048, relu=True)
    # weight gradients
    dq = torch.randn(b * n, 768, device=dev, generator=g).to(BF16); y = torch.randn(b * n, 256, device=dev, generator=g).to(BF16)
    fs.wgrad_(torch.zeros(768, 256, device=dev), dq, y)
    dk = torch.randn(b * m, 256, device=dev, generator=g).to(BF16)
    fs.wgrad_(torch.zeros(256, 64, device=dev), dk, mem.view(b * m, 64))
    torch.cuda.synchronize()
print("done")
