"""Per-kernel timing of the attention backward through `parts` (1 Delta, 2 dV, 4 dK, 8 dQ; 6 = key side), real-call mode."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ops
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0")
for shape in sys.argv[1:]:
    b, n, m = (int(x) for x in shape.split(","))
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn(b, n, 256, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    k = (torch.randn(b, m, 256, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    grid = int(round(math.sqrt(n)))
    kw = dict(table=compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev), n_rope_k=(m // n) * n, grad_dtype=torch.bfloat16)
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0)
    delta = torch.empty(b, n, device=dev)
    ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, parts=1, delta=delta, **kw)
    res = {}
    for name, parts in (("delta", 1), ("dV", 2), ("dK", 4), ("key side (2|4)", 6), ("dQ", 8)):
        db = tuple(torch.zeros(256, device=dev) for _ in range(3))
        for _ in range(3):
            ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, parts=parts, delta=delta, dbias=db, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, parts=parts, delta=delta, dbias=db, **kw)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10 * 1e3
    print(f"B={b} N={n} M={m}: " + "  ".join(f"{k} {v:.1f} us" for k, v in res.items()), flush=True)
