"""Micro-benchmark of the attention kernels through the C ABI (GPU box).  CUDA-event timing.
usage: python scripts/attn_kernel_bench.py [shape ...]   shape = B,N,M   (default: a few BASELINE shapes)"""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ops, _lib

def bench(b, n, m, iters=int(os.environ.get('BENCH_ITERS', '10')), bwd=True):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn(b, n, 256, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    k = (torch.randn(b, m, 256, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    scale = 1 / 16.0
    # ATTN_BENCH_REAL=1: call the backward the way the fused stack does (bf16 gradients, conjugate RoPE and bias
    # gradients fused into the epilogues); default: plain fp32 gradients
    kw = {}
    if os.environ.get("ATTN_BENCH_REAL"):
        from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
        grid = int(round(math.sqrt(n)))
        if grid * grid == n and (m // n) * n <= m:
            kw = dict(table=compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev), n_rope_k=(m // n) * n,
                      grad_dtype=torch.bfloat16, dbias=tuple(torch.zeros(256, device=dev) for _ in range(3)))
    for _ in range(int(os.environ.get('BENCH_WARMUP', '3'))):
        o, o32, lse = ops.attn_fwd(q, k, v, scale)
        if bwd: ops.attn_bwd(q, k, v, None, o32, do, lse, scale, **kw)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(iters): o, o32, lse = ops.attn_fwd(q, k, v, scale)
    e[1].record()
    if bwd:
        for _ in range(iters): ops.attn_bwd(q, k, v, None, o32, do, lse, scale, **kw)
    e[2].record()
    torch.cuda.synchronize()
    tf, tb = e[0].elapsed_time(e[1]) / iters, e[1].elapsed_time(e[2]) / iters
    fl = 4.0 * b * n * m * 256
    ns = _lib.load().sam2b200_attn_default_nsplit(b, n, m)
    print(f"B={b} N={n} M={m} nsplit={ns}: fwd {tf*1e3:8.1f} us {fl/tf/1e9:7.1f} TF/s | bwd {tb*1e3:8.1f} us {2.5*fl/max(tb,1e-9)/1e9:7.1f} TF/s (2.5x fwd flops)", flush=True)

if __name__ == "__main__":
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [
        (56, 576, 4060), (56, 576, 576), (13, 1024, 7196), (4, 4096, 28736), (4, 4096, 4096), (1, 576, 4060)]
    for s in shapes:
        bench(*s)
