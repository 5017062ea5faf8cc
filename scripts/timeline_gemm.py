"""Per-CTA %globaltimer phase timeline of sam2b200_gemm_ex (csrc/gemm.cu) at the cfg2 shapes.  usage (GPU box): python scripts/timeline_gemm.py [B N]"""
import os, sys, math, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs, _lib
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
b, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (56, 576)
dev = torch.device("cuda:0"); BF16 = torch.bfloat16
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
r = b * n
table = compute_axial_cis(dim=256, end_x=int(math.sqrt(n)), end_y=int(math.sqrt(n))).to(dev)
y = torch.randn(r, 256, device=dev, generator=g).to(BF16)
h = torch.randn(r, 2048, device=dev, generator=g).to(BF16)
cases = {
    "qkv head  K=256 Nout=768 rope (BN=128)": lambda: fs.gemm_ex(y, wqkv, 3, 256, bias=b768, table=table, rope_outs=2, rows_per_item=n, n_rope_rows=n),
    "mlp head  K=256 Nout=2048 relu (BN=256)": lambda: fs.gemm_ex(y, w1, 1, 2048, bias=b2048, relu=True),
    "d out_proj K=256 Nout=256 NN (BN=256)": lambda: fs.gemm(y, wo, nn=True),
    "linear2   K=2048 Nout=256 (BN=256, streamed weights)": lambda: fs.gemm(h, w2, bias=b256),
}
wqkv = (torch.randn(768, 256, device=dev, generator=g) / 16).to(BF16); b768 = torch.zeros(768, device=dev)
w1 = (torch.randn(2048, 256, device=dev, generator=g) / 16).to(BF16); b2048 = torch.zeros(2048, device=dev)
wo = (torch.randn(256, 256, device=dev, generator=g) / 16).to(BF16)
w2 = (torch.randn(256, 2048, device=dev, generator=g) / 45).to(BF16); b256 = torch.zeros(256, device=dev)
def med(v): return f"{np.median(v):7.0f} [{np.percentile(v, 10):6.0f}..{np.percentile(v, 90):6.0f}]"
print(f"# R = {r}; ns (%globaltimer), median [p10..p90] over the CTAs of one launch")
for name, fn in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    buf = torch.zeros(32 * 4096, dtype=torch.int64, device=dev)
    lib.sam2b200_gemm_debug_timeline(buf.data_ptr(), buf.numel())
    fn(); torch.cuda.synchronize()
    used = lib.sam2b200_gemm_debug_timeline(None, 0)
    t = buf[:used].cpu().numpy().reshape(-1, 32)
    t0 = t[:, 0].min()
    print(f"{name}: {len(t)} CTAs, tiles per CTA {int(t[:, 31].min())}..{int(t[:, 31].max())}, span {(t[:, 30].max() - t0) / 1e3:.1f} us; start skew {med(t[:, 0] - t0)}")
    nt = int(t[:, 31].min())
    for i in range(min(nt, 3)):
        free, issued, done, stored = t[:, 2 + 4 * i], t[:, 3 + 4 * i], t[:, 4 + 4 * i], t[:, 5 + 4 * i]
        line = f"  tile {i}: acc free @ {med(free - t[:, 0])} | loads landed + MMAs issued +{med(issued - free)} | MMAs complete +{med(done - issued)} | epilogue (warp 0) {med(stored - done)}"
        if i + 1 < min(nt, 4): line += f" | period to next tile {med(t[:, 4 + 4 * (i + 1)] - done)}"
        print(line)
