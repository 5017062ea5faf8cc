"""Per-CTA phase timeline of the attention kernels (GPU box).  usage: python scripts/timeline.py B N M"""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ops, _lib
b, n, m = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (56, 576, 4060)
dev = torch.device("cuda:0")
lib = _lib.load()
q = torch.randn(b, n, 256, device=dev).to(torch.bfloat16); k = torch.randn(b, m, 256, device=dev).to(torch.bfloat16)
v = torch.randn(b, m, 256, device=dev).to(torch.bfloat16); do = torch.randn(b, n, 256, device=dev).to(torch.bfloat16)
for _ in range(2):
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16, 1)
    ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16)
torch.cuda.synchronize()
buf = torch.zeros(8 * 20000, dtype=torch.int64, device=dev)
lib.sam2b200_debug_set_timeline(buf.data_ptr(), buf.numel())
o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16, 1)
ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16)
torch.cuda.synchronize()
used = lib.sam2b200_debug_set_timeline(None, 0)
t = buf[:used].cpu().numpy().reshape(-1, 8)
gq, gk = (n + 127) // 128 * b, (m + 127) // 128 * b
names = [("fwd", gq), ("dV", gk), ("dK", gk), ("dQ", gq)]
off = 0
for name, g in names:
    e = t[off:off + g]; off += g
    t0 = e[:, 1].min()
    tot = e[:, 6].max() - t0
    setup = (e[:, 2] - e[:, 1]); oper = (e[:, 3] - e[:, 2]); first = (e[:, 4] - e[:, 3]); loop = (e[:, 5] - e[:, 4]); epi = (e[:, 6] - e[:, 5])
    print(f"{name}: {g} CTAs, kernel span {tot/1e3:.1f} us, tiles/CTA {int(e[0,7])}")
    print(f"   per-CTA ns (median): setup {np.median(setup):.0f}  operand->TMEM {np.median(oper):.0f}  first scores {np.median(first):.0f}  "
          f"main loop {np.median(loop):.0f} ({np.median(loop)/max(e[0,7]-1,1):.0f}/tile)  epilogue {np.median(epi):.0f}  total {np.median(e[:,6]-e[:,1]):.0f}")
    # gap between consecutive CTAs on the same SM
    gaps = []
    for sm in np.unique(e[:, 0]):
        ee = e[e[:, 0] == sm]; ee = ee[np.argsort(ee[:, 1])]
        gaps += list(ee[1:, 1] - ee[:-1, 6])
    if gaps: print(f"   gap between CTAs on an SM: median {np.median(gaps):.0f} ns, p90 {np.percentile(gaps,90):.0f} ns; CTAs per SM max {max(np.bincount(e[:,0].astype(int)))}")
