"""sam2b200_wgrad (split-K tcgen05, in-place fp32 accumulation) vs cuBLAS addmm(out_dtype=fp32) for the weight-gradient
shapes of one layer at cfg2 (R = 56 x 576, memory rows 56 x 4060).  Device time: 10 calls per CUDA graph, replayed."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs
dev = torch.device("cuda:0")
BF16, F32 = torch.bfloat16, torch.float32
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(it): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * it) * 1e3
b, n, m = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (56, 576, 4060)
r, rm = b * n, b * m
gen = torch.Generator(device="cuda").manual_seed(0)
tot = [0.0, 0.0]
for name, rows, mo, no in (("l2.w  [256 x 2048]", r, 256, 2048), ("l1.w  [2048 x 256]", r, 2048, 256), ("qkv.w [768 x 256]", r, 768, 256),
                           ("sa.o.w [256 x 256]", r, 256, 256), ("ca.q.w [256 x 256]", r, 256, 256), ("ca.k.w [256 x 64] (memory rows)", rm, 256, 64),
                           ("fold G [256 x 64]", r, 256, 64)):
    a = torch.randn(rows, mo, device=dev, generator=gen).to(BF16)
    x = torch.randn(rows, no, device=dev, generator=gen).to(BF16)
    c = torch.zeros(mo, no, device=dev)
    t_own = timeit(lambda: fs.wgrad_(c, a, x))
    t_lib = timeit(lambda: torch.addmm(c, a.t(), x, out_dtype=F32, out=c))
    tot[0] += t_own; tot[1] += t_lib
    byt = (a.numel() + x.numel()) * 2
    print(f"{name:34s} R={rows:6d}: wgrad {t_own:7.1f} us ({byt / t_own / 1e3:5.0f} GB/s, {2.0 * rows * mo * no / t_own / 1e6:5.0f} TF/s) | cuBLAS {t_lib:7.1f} us", flush=True)
print(f"sum: wgrad {tot[0]:.1f} us | cuBLAS {tot[1]:.1f} us")
