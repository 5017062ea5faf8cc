"""Micro-benchmark (GPU box): sam2b200_proj_rope (projection GEMM with bias + RoPE in the epilogue) vs cuBLAS addmm + rope pass."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs
from sam2_video_training_b200.ops import rope_apply
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis

def bench(name, rows, k, n_out, rope_outs, b, length, n_rope, grid=24, iters=20):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    xs = [torch.randn(rows, k, device=dev, generator=g).to(torch.bfloat16) for _ in range(3)]
    w = (torch.randn(256 * n_out, k, device=dev, generator=g) / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn(256 * n_out, device=dev, generator=g).to(torch.bfloat16)
    def ours(i): return fs.proj_rope(xs[i % 3], w, bias, n_out, table if rope_outs else None, rope_outs, length, n_rope)
    def theirs(i):
        outs = []
        for j in range(n_out):
            y = torch.addmm(bias[256 * j:256 * (j + 1)], xs[i % 3], w[256 * j:256 * (j + 1)].t())
            if j < rope_outs: y = rope_apply(y.view(b, length, 256), table, n_rope)
            outs.append(y)
        return outs
    def timeit(fn):
        for i in range(3): fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    t0, t1 = timeit(ours), timeit(theirs)
    byts = (rows * k + rows * 256 * n_out) * 2
    print(f"{name:28s} R={rows} K={k} Nout={256*n_out}: proj_rope {t0*1e3:6.1f} us ({byts/t0/1e6:5.0f} GB/s) | addmm + rope {t1*1e3:6.1f} us | x{t1/t0:.2f}", flush=True)

if __name__ == "__main__":
    bench("self q|k|v (2 rotated)", 56 * 576, 256, 3, 2, 56, 576, 576)
    bench("cross q (rotated)", 56 * 576, 256, 1, 1, 56, 576, 576)
    bench("memory keys (rotated)", 56 * 4060, 64, 1, 1, 56, 4060, 4032)
    bench("memory values", 56 * 4060, 64, 1, 0, 56, 4060, 0)
    bench("cfg4 self q|k|v", 4 * 4096, 256, 3, 2, 4, 4096, 4096, grid=64)
    bench("cfg4 memory keys", 4 * 28736, 64, 1, 1, 4, 28736, 28672, grid=64)
