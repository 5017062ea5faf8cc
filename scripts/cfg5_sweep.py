"""BASELINE.json cfg5: MemoryAttention stack fwd+bwd at N = 1024 (512 px), memory frames n_f = 1..16
(M = n_f * 1024 + 4 * min(n_f, 16) pointer tokens), B objects.  CUDA events per iteration, >= 10 warm-ups, 50 timed
iterations, median + min; algorithmic FLOPs per SURVEY.md section 8d (attention core 3.5 x fwd, linears 3 x fwd)."""
import os, sys, json, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import ddp
from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
from sam2_video_training_b200.graphs import GraphedMemoryAttention

dev = torch.device("cuda:0")
B = int(os.environ.get("CFG5_B", "13"))
N, d, dm, ff, L = 1024, 256, 64, 2048, 4
model = build_memory_attention(dropout=0.0).to(dev).train()
ddp.attach_grad_bucket(model)
fwd = GraphedMemoryAttention(model)
g = torch.Generator(device="cuda").manual_seed(5)
curr_pos = torch.randn(N, B, d, device=dev, generator=g) * 0.7
for nf in (1, 2, 4, 8, 16):
    P = 4 * min(nf, 16)
    M = nf * N + P
    sets = []
    for _ in range(3):
        sets.append((torch.randn(N, B, d, device=dev, generator=g), torch.randn(M, B, dm, device=dev, generator=g),
                     (torch.randn(M, B, dm, device=dev, generator=g) * 0.7).requires_grad_(True),
                     torch.randn(N, B, d, device=dev, generator=g)))
    def step(k):
        curr, mem, pos, go = sets[k % 3]
        out = fwd(curr, mem, curr_pos, pos, P)
        out.backward(go)
        pos.grad = None
    for k in range(10):
        step(k)
    torch.cuda.synchronize()
    ts = []
    for k in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(k); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    core = 4.0 * N * N * d + 4.0 * N * M * d
    lin = 4 * (2.0 * N * d * d) + 2 * (2.0 * N * d * d) + 2 * (2.0 * M * dm * d) + 2 * (2.0 * N * d * ff)
    fl = L * B * (3.5 * core + 3.0 * lin)
    med, mn = statistics.median(ts), min(ts)
    print(json.dumps({"n_f": nf, "B": B, "N": N, "M": M, "ms_median": round(med, 3), "ms_min": round(mn, 3),
                      "gflop_per_object_frame": round(fl / B / 1e9, 1), "tflops_median": round(fl / med / 1e9, 1),
                      "attention_core_share_of_flops": round(L * B * 3.5 * core / fl, 3),
                      "object_frames_per_s": round(B / (med * 1e-3), 1)}), flush=True)
    model._sam2b200_grad_bucket.zero()
