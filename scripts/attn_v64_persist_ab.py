"""A/B of the raw-memory cross-attention backward: one CTA per (row block, object) item (variant key 0 = 2) vs resident CTAs
(three_gemm_v64_persistent_kernel, default).  bf16 gradients, conjugate RoPE in the epilogue, no bias gradients (they come
from sam2b200_wgrad), CUDA events, 20 launches after 5 warm-ups.  Usage: python scripts/attn_v64_persist_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam2_video_training_b200 import _lib, ops
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis

lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for b, grid, m in ((56, 24, 580), (56, 24, 1160), (56, 24, 2320), (56, 24, 4060), (13, 32, 7196), (4, 64, 28736)):
    n = grid * grid
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    o64, o32, lse, _ = ops.attn_fwd_v64(q, k, mem, 1 / 16.0)
    delta = (do64.float() * o32).sum(-1)
    kw = dict(table=table, n_rope_k=(m // n) * n, grad_dtype=torch.bfloat16)
    for variant, name in ((2, "one CTA per item"), (0, "resident CTAs   ")):
        lib.sam2b200_debug_set_variant(0, variant)
        tk = timeit(lambda: ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16.0, parts=4, **kw))
        tq = timeit(lambda: ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16.0, parts=8, **kw))
        fl = 2.0 * b * n * m * (256 + 64 + 256) * 2
        print(f"B={b} N={n} M={m} {name}: dK {tk:7.1f} us  dQ {tq:7.1f} us  bwd {tk + tq:7.1f} us = {fl / (tk + tq) / 1e6:5.0f} TF/s algorithmic", flush=True)
lib.sam2b200_debug_set_variant(0, 0)
