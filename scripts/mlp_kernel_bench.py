"""Micro-benchmark (GPU box): sam2b200_mlp_dh (GEMM + fused ReLU backward) vs torch.mm + the separate mask / bias-gradient pass.
Algorithmic bytes: read dm (R x 256) + h (R x F), write dh (R x F), bf16."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs

def bench(rows, f=2048, iters=20):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    dm = torch.randn(rows, 256, device=dev, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(256, f, device=dev, generator=g) / 16).to(torch.bfloat16)
    hs = [torch.relu(torch.randn(rows, f, device=dev, generator=g)).to(torch.bfloat16) for _ in range(3)]   # rotate: > L2
    bias = torch.zeros(f, device=dev)
    def ours(i): return fs.mlp_dh(dm, w2, hs[i % 3], 1.0)
    def theirs(i):
        dh = torch.mm(dm, w2); fs.relu_bwd_colsum_(dh, hs[i % 3], bias); return dh
    def timeit(fn):
        for i in range(3): fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    t0, t1 = timeit(ours), timeit(theirs)
    byts = (rows * 256 + 2 * rows * f) * 2
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(f"R={rows} F={f}: mlp_dh {t0*1e3:6.1f} us ({byts/t0/1e6:5.0f} GB/s = {byts/t0/1e6/pk:5.1%} of {pk:.0f}; {2.0*rows*256*f/t0/1e9:6.1f} TF/s) | "
          f"torch.mm + mask pass {t1*1e3:6.1f} us | x{t1/t0:.2f}", flush=True)

if __name__ == "__main__":
    for r in [int(a) for a in sys.argv[1:]] or [32256, 13312, 16384, 576]:
        bench(r)
