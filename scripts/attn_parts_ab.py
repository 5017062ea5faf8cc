"""A/B of the gradient-epilogue rotation-table addressing (axial rows vs full rows) on every backward kernel family,
cfg2 shapes.  CUDA events, bf16 gradients + fused conjugate RoPE + bias gradients (real-call mode)."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ops, _lib
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0")
lib = _lib.load()
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for shape in sys.argv[1:] or ["56,576,4060", "56,576,576", "13,1024,7196", "13,1024,1024"]:
    b, n, m = (int(x) for x in shape.split(","))
    g = torch.Generator(device="cuda").manual_seed(1)
    q, do = (torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
    k, v = (torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
    grid = int(round(math.sqrt(n)))
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    nr = (m // n) * n
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0)
    o64, o64_32, lse64, _ = ops.attn_fwd_v64(q, k, mem, 1 / 16.0)
    delta = (do64.float() * o64_32).sum(-1)
    db = tuple(torch.zeros(256, device=dev) for _ in range(3))
    for variant, name in ((1, "full rows "), (0, "axial rows")):
        lib.sam2b200_debug_set_variant(1, variant)
        kw = dict(table=table, n_rope_k=nr, grad_dtype=torch.bfloat16)
        t_dq = timeit(lambda: ops.attn_bwd_v64(q, k, mem, do64, lse64, delta, 1 / 16.0, dbias=(db[0], None), parts=8, **kw))
        t_dk = timeit(lambda: ops.attn_bwd_v64(q, k, mem, do64, lse64, delta, 1 / 16.0, dbias=(None, db[1]), parts=4, **kw))
        t_256 = timeit(lambda: ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, dbias=db, **kw))
        print(f"B={b} N={n} M={m} {name}: raw-memory dQ {t_dq:7.1f} dK {t_dk:7.1f} us | 256-d backward (dV+dK+dQ) {t_256:7.1f} us", flush=True)
    lib.sam2b200_debug_set_variant(1, 0)
