"""Pre-norm block heads: the one-kernel sam2b200_ln_proj vs ln_fwd + sam2b200_gemm_ex (resident-CTA GEMM with the bias / RoPE / ReLU epilogue),
plus sam2b200_gemm vs cuBLAS for the other dense GEMMs of a layer.  Device time: 12 calls per CUDA graph, inputs rotated over 4 sets."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs
from sam2_video_training_b200.ops import rope_apply
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0")
BF16 = torch.bfloat16
def timeit(fn, it=12):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(it): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * it) * 1e3
shapes = ((56, 576), (13, 1024), (4, 4096)) if len(sys.argv) < 2 else ((int(sys.argv[1]), int(sys.argv[2])),)
for b, n in shapes:
    r = b * n
    g = torch.Generator(device="cuda").manual_seed(0)
    nset = 4
    xs = [torch.randn(r, 256, device=dev, generator=g) for _ in range(nset)]
    rs = [torch.randn(r, 256, device=dev, generator=g).to(BF16) for _ in range(nset)]
    gamma, beta = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    table = compute_axial_cis(dim=256, end_x=int(math.sqrt(n)), end_y=int(math.sqrt(n))).to(dev)
    for mode, nout, n_out, width, ropes, relu in (("qkv", 768, 3, 256, 2, False), ("q  ", 256, 1, 256, 1, False), ("mlp", 2048, 1, 2048, 0, True)):
        w = (torch.randn(nout, 256, device=dev, generator=g) / 16).to(BF16)
        b32 = torch.zeros(nout, device=dev); b16 = b32.to(BF16)
        kw = dict(out_width=width, table=table if ropes else None, rope_outs=ropes, rows_per_item=n, n_rope_rows=n, relu=relu)
        tf = timeit(lambda i: fs.ln_proj(xs[i % nset], rs[i % nset], gamma, beta, w, b16, n_out, **kw))
        ts = timeit(lambda i: fs.ln_then_proj(xs[i % nset], rs[i % nset], gamma, beta, w, b32, n_out, **kw))
        tl = timeit(lambda i: fs.ln_fwd(xs[i % nset], rs[i % nset], gamma, beta))
        ys = [fs.ln_fwd(xs[i], rs[i], gamma, beta)[0] for i in range(nset)]
        tg = timeit(lambda i: fs.gemm_ex(ys[i % nset], w, n_out, width, bias=b32, table=kw["table"], rope_outs=ropes, rows_per_item=n, n_rope_rows=n, relu=relu))
        mb = (r * 256 * 2 + r * nout * 2 + nout * 512) / 1e6
        print(f"R={r:6d} {mode} Nout={nout:4d}: ln_proj {tf:6.1f} us | ln_fwd + gemm_ex {ts:6.1f} us (ln_fwd {tl:5.1f}, gemm_ex {tg:5.1f} us = {mb / tg:5.2f} TB/s of {mb:.0f} MB)", flush=True)
    # the other dense GEMMs of a layer: sam2b200_gemm vs cuBLAS
    for name, k, no, nn in (("linear2 fwd  [R,2048]x[256,2048]^T", 2048, 256, False), ("d linear1    [R,2048]x[2048,256]", 2048, 256, True),
                            ("d qkv        [R,768]x[768,256]", 768, 256, True), ("d out_proj   [R,256]x[256,256]", 256, 256, True),
                            ("d folded     [R,256]x[256,64]", 256, 64, True)):
        a = [torch.randn(r, k, device=dev, generator=g).to(BF16) for _ in range(nset)]
        w = (torch.randn((k, no) if nn else (no, k), device=dev, generator=g) / 16).to(BF16)
        bias32 = torch.zeros(no, device=dev); bias16 = bias32.to(BF16)
        to = timeit(lambda i: fs.gemm(a[i % nset], w, nn=nn, bias=None if nn else bias32))
        tc = timeit(lambda i: (torch.mm(a[i % nset], w) if nn else torch.addmm(bias16, a[i % nset], w.t())))
        mb = (r * k * 2 + r * no * 2 + k * no * 2) / 1e6
        print(f"R={r:6d} {name:36s}: sam2b200_gemm {to:6.1f} us ({mb / to:5.2f} TB/s of {mb:.0f} MB) | cuBLAS {tc:6.1f} us", flush=True)
    # memory-key projection + RoPE (K = 64), M = 7 frames + 16 pointers x 4 tokens
    m = 7 * n + 64
    mem = [torch.randn(b * m, 64, device=dev, generator=g).to(BF16) for _ in range(2)]
    wk = (torch.randn(256, 64, device=dev, generator=g) / 8).to(BF16); bk32 = torch.zeros(256, device=dev); bk16 = bk32.to(BF16)
    to = timeit(lambda i: fs.gemm(mem[i % 2], wk, bias=bk32, table=table, rows_per_item=m, n_rope_rows=7 * n))
    tc = timeit(lambda i: rope_apply(torch.addmm(bk16, mem[i % 2], wk.t()).view(b, m, 256), table, 7 * n))
    mb = (b * m * 64 * 2 + b * m * 512) / 1e6
    print(f"R={b * m:6d} memory keys + RoPE [B M,64]x[256,64]^T   : sam2b200_gemm {to:6.1f} us ({mb / to:5.2f} TB/s of {mb:.0f} MB) | cuBLAS + rope pass {tc:6.1f} us", flush=True)
