"""Cross-attention core: 256-d value path (attn_fwd / attn_bwd incl. dV) vs the raw-memory path (attn_fwd_v64 / attn_bwd_v64),
real-call mode (bf16 gradients, conjugate RoPE + bias gradients in the epilogues).  CUDA events."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ops
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0")
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for shape in sys.argv[1:] or ["56,576,4060", "56,576,1160", "13,1024,7196", "4,4096,28736"]:
    b, n, m = (int(x) for x in shape.split(","))
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn(b, n, 256, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    k = (torch.randn(b, m, 256, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
    grid = int(round(math.sqrt(n)))
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    nr = (m // n) * n
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0)
    o64, o64_32, lse64, _ = ops.attn_fwd_v64(q, k, mem, 1 / 16.0)
    delta = torch.zeros(b, n, device=dev)
    db = tuple(torch.zeros(256, device=dev) for _ in range(3))
    t_f = timeit(lambda: ops.attn_fwd(q, k, v, 1 / 16.0))
    t_f64 = timeit(lambda: ops.attn_fwd_v64(q, k, mem, 1 / 16.0))
    t_b = timeit(lambda: ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=nr, grad_dtype=torch.bfloat16, dbias=db))
    t_b64 = timeit(lambda: ops.attn_bwd_v64(q, k, mem, do64, lse64, delta, 1 / 16.0, table=table, n_rope_k=nr, grad_dtype=torch.bfloat16, dbias=db[:2]))
    t_dq = timeit(lambda: ops.attn_bwd_v64(q, k, mem, do64, lse64, delta, 1 / 16.0, table=table, n_rope_k=nr, grad_dtype=torch.bfloat16, dbias=db[:2], parts=8))
    fl = 4.0 * b * n * m * 256
    print(f"B={b} N={n} M={m}: fwd {t_f:7.1f} -> {t_f64:7.1f} us | bwd {t_b:7.1f} -> {t_b64:7.1f} us (dQ {t_dq:.1f}, dK {t_b64 - t_dq:.1f}) | "
          f"fwd+bwd {t_f + t_b:7.1f} -> {t_f64 + t_b64:7.1f} us = {3.5 * fl / (t_f64 + t_b64) / 1e6:.0f} TF/s algorithmic (was {3.5 * fl / (t_f + t_b) / 1e6:.0f})", flush=True)
