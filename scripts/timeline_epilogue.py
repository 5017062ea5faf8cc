"""Inside-the-epilogue timeline of three_gemm_v64_kernel<DK|DQ> (library built with -DSAM2B200_EPI_TIMELINE: slots 2, 3, 4, 7 are
re-purposed as 'after chunk 0 / chunk 1 / chunk 3 / TMA store read' stamps of warp 0).  Four variants: rotation on/off, bias gradient on/off.
usage (GPU box): SAM2B200_LIB=sam2_video_training_b200/build/libsam2b200_epi.so python scripts/timeline_epilogue.py [B N M]"""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import _lib
if os.environ.get("SAM2B200_LIB"):
    _lib.LIB_PATH = os.path.join(ROOT, os.environ["SAM2B200_LIB"])
from sam2_video_training_b200 import ops
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
table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
db = torch.zeros(256, device=dev)
print(f"# B={b} N={n} M={m}; warp 0 of each CTA, ns (%globaltimer), median [p10 .. p90]")
for rot in (True, False):
    for bias in (True, False):
        kw = dict(table=table if rot else None, n_rope_k=(m // n) * n if rot else 0, grad_dtype=torch.bfloat16)
        dbk, dbq = ((None, db), (db, None)) if bias else ((None, None), (None, None))
        for _ in range(3):
            ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=4, dbias=dbk, **kw)
            ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=8, dbias=dbq, **kw)
        torch.cuda.synchronize()
        buf = torch.zeros(8 * 40000, dtype=torch.int64, device=dev)
        lib.sam2b200_debug_set_timeline(buf.data_ptr(), buf.numel())
        ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=4, dbias=dbk, **kw)
        ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16, parts=8, dbias=dbq, **kw)
        torch.cuda.synchronize()
        used = lib.sam2b200_debug_set_timeline(None, 0)
        t = buf[:used].cpu().numpy().reshape(-1, 8)
        gk, gq = (m + 127) // 128 * b, (n + 127) // 128 * b
        off = 0
        for name, gsz in (("dK", gk), ("dQ", gq)):
            e = t[off:off + gsz]; off += gsz
            ph = {"start -> last MMA done": e[:, 5] - e[:, 1], "chunk 0": e[:, 2] - e[:, 5], "chunk 1 (+ first TMA store)": e[:, 3] - e[:, 2],
                  "chunks 2, 3 (+ second store)": e[:, 4] - e[:, 3], "TMA store has read SMEM": e[:, 7] - e[:, 4], "CTA barrier (other warps)": e[:, 6] - e[:, 7],
                  "epilogue total": e[:, 6] - e[:, 5]}
            print(f"{name} rotation={int(rot)} bias_grad={int(bias)}: span {(e[:, 6].max() - e[:, 1].min()) / 1e3:.1f} us; " +
                  "; ".join(f"{kn} {np.median(v):.0f} [{np.percentile(v, 10):.0f}..{np.percentile(v, 90):.0f}]" for kn, v in ph.items()), flush=True)
