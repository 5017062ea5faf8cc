"""A/B of the raw-memory cross-attention forward: single-stream kernel (default) vs two-stream kernel (variant key 2 = 1).
cfg2 / cfg3 / cfg4 shapes, CUDA-event timed on the launching stream, 20 launches after 5 warm-ups each.
Usage: python scripts/attn_fwd_v64_ab.py > gpurun_out/attn_fwd_v64_ab.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam2_video_training_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
for name, b, n, m in (("cfg2", 56, 576, 4060), ("cfg3", 24, 576, 4060), ("cfg4", 8, 4096, 28736), ("early", 56, 576, 580)):
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = (0.5 * torch.randn(b, m, 256, device=dev, generator=g)).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    w16 = (0.1 * torch.randn(256, 64, device=dev, generator=g)).to(torch.bfloat16)
    bias = torch.randn(256, device=dev, generator=g)
    for proj in (False, True):
        for variant in (0, 1):
            lib.sam2b200_debug_set_variant(2, variant)
            fn = (lambda: ops.attn_fwd_v64_proj(q, k, mem, w16, bias, None, 1 / 16.0)) if proj else (lambda: ops.attn_fwd_v64(q, k, mem, 1 / 16.0))
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            flops = 2.0 * b * n * m * (256 + 64)
            print(f"{name} proj={int(proj)} kernel={'two-stream' if variant == 1 else 'single'}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
lib.sam2b200_debug_set_variant(2, 0)
