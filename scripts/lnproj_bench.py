"""sam2b200_ln_proj vs the op sequence it replaces (ln_fwd + cuBLAS addmm + RoPE pass / ReLU epilogue), cfg2 / cfg3 / cfg4 row counts.
Device time: 12 calls captured into a CUDA graph and replayed (inputs rotated over 4 sets)."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs
from sam2_video_training_b200.ops import rope_apply
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0")
BF16 = torch.bfloat16
def timeit(fn, it=12):
    """Device time per call: `it` calls captured into one CUDA graph (no host launch overhead), replayed 5 times."""
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
for b, n in ((56, 576), (13, 1024), (4, 4096)):
    r = b * n
    g = torch.Generator(device="cuda").manual_seed(0)
    nset = 4
    xs = [torch.randn(r, 256, device=dev, generator=g) for _ in range(nset)]
    rs = [torch.randn(r, 256, device=dev, generator=g).to(BF16) for _ in range(nset)]
    gamma, beta = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    grid = int(math.sqrt(n))
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    for mode, nout, n_out, width, ropes, relu in (("qkv", 768, 3, 256, 2, False), ("q  ", 256, 1, 256, 1, False), ("mlp", 2048, 1, 2048, 0, True)):
        w = (torch.randn(nout, 256, device=dev, generator=g) / 16).to(BF16)
        bias = torch.zeros(nout, device=dev, dtype=BF16)
        def fused(i):
            fs.ln_proj(xs[i % nset], rs[i % nset], gamma, beta, w, bias, n_out, out_width=width, table=table if ropes else None, rope_outs=ropes,
                       rows_per_item=n, n_rope_rows=n, relu=relu)
        def seq(i):
            y, x_new, mean, rstd = fs.ln_fwd(xs[i % nset], rs[i % nset], gamma, beta)
            if relu:
                torch._addmm_activation(bias, y, w.t(), use_gelu=False)
            else:
                for j in range(n_out):
                    o = torch.addmm(bias[j * 256:(j + 1) * 256], y, w[j * 256:(j + 1) * 256].t())
                    if j < ropes: rope_apply(o.view(b, n, 256), table, n)
        def ln_only(i):
            fs.ln_fwd(xs[i % nset], rs[i % nset], gamma, beta)
        tf, ts, tl = timeit(fused), timeit(seq), timeit(ln_only)
        fl = 2.0 * r * 256 * nout
        print(f"R={r:6d} {mode} Nout={nout:4d}: ln_proj {tf:7.1f} us ({fl / tf / 1e6:5.0f} TF/s) | ln_fwd + cuBLAS (+ rope) {ts:7.1f} us (ln_fwd alone {tl:.1f})", flush=True)
