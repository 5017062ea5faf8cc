"""Per-CTA %globaltimer phase timeline of the raw-memory cross-attention backward kernels (three_gemm_v64_kernel<DQ|DK>).
usage (GPU box): python scripts/timeline_v64.py [B N M]        default: cfg2 cross shape 56 576 4060"""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ops, _lib
b, n, m = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (56, 576, 4060)
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16); k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16); do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
o64, o32, lse, _ = ops.attn_fwd_v64(q, k, mem, 1 / 16)
delta = (do64.float() * o32).sum(-1)
grid = int(n ** 0.5)
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev) if grid * grid == n else None
kw = dict(table=table, n_rope_k=(m // n) * n if table is not None else 0, grad_dtype=torch.bfloat16)
db = torch.zeros(256, device=dev)
for _ in range(3):
    ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=4, dbias=(None, db), **kw)
    ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=8, dbias=(db, None), **kw)
torch.cuda.synchronize()
buf = torch.zeros(8 * 40000, dtype=torch.int64, device=dev)
lib.sam2b200_debug_set_timeline(buf.data_ptr(), buf.numel())
ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=4, dbias=(None, db), **kw)
ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=8, dbias=(db, None), **kw)
torch.cuda.synchronize()
used = lib.sam2b200_debug_set_timeline(None, 0)
t = buf[:used].cpu().numpy().reshape(-1, 8)
gk, gq = (m + 127) // 128 * b, (n + 127) // 128 * b
off = 0
print(f"# B={b} N={n} M={m}; per-CTA phases in ns (%globaltimer), median [p10 .. p90]")
for name, gsz in (("dK three_gemm_v64_kernel<1>", gk), ("dQ three_gemm_v64_kernel<0>", gq)):
    e = t[off:off + gsz]; off += gsz
    if len(e) == 0: continue
    nt = int(e[0, 7])
    t0 = e[:, 1].min()
    span = e[:, 6].max() - t0
    ph = {"setup (barriers, TMEM alloc)": e[:, 2] - e[:, 1], "operands landed (TMA A1+A2)": e[:, 3] - e[:, 2],
          "first S ready (fill)": e[:, 4] - e[:, 3], "main loop": e[:, 5] - e[:, 4], "epilogue (drain+store)": e[:, 6] - e[:, 5],
          "CTA total": e[:, 6] - e[:, 1]}
    print(f"{name}: {gsz} CTAs x {nt} tiles, kernel span {span/1e3:.1f} us")
    for kname, v in ph.items():
        print(f"   {kname:32s} {np.median(v):8.0f} [{np.percentile(v,10):.0f} .. {np.percentile(v,90):.0f}]" + (f"   = {np.median(v)/max(nt-1,1):.0f} / tile" if kname == "main loop" else ""))
    gaps = []
    per_sm = []
    for sm in np.unique(e[:, 0]):
        ee = e[e[:, 0] == sm]; ee = ee[np.argsort(ee[:, 1])]
        gaps += list(ee[1:, 1] - ee[:-1, 6]); per_sm.append(len(ee))
    if gaps: print(f"   gap between consecutive CTAs on an SM: median {np.median(gaps):.0f} ns, p90 {np.percentile(gaps,90):.0f}; CTAs per SM {min(per_sm)}..{max(per_sm)}")
    mma_ns = nt * 1152 / 1.9
    print(f"   MMA-only time per CTA at 1152 clk/tile @1.9 GHz: {mma_ns:.0f} ns -> tensor-pipe bound of this schedule {mma_ns/np.median(ph['CTA total'])*100:.0f} % of CTA time")
