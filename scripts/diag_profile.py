"""Diagnostic: torch-profiler kernel table for one bench step (GPU box)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sam2_video_training_b200 import ddp
from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_endovis18_384px_T10_7obj_x8clips"]
dev = torch.device("cuda:0")
model = build_memory_attention(dropout=0.0).to(dev).train()
crit = MultiStepMultiMasksAndIous(dict(bench.LOSS_W), supervise_all_iou=True, iou_use_l1_loss=True, check_valid=False)
opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True)
ddp.attach_grad_bucket(model)
host = bench.make_host_inputs(wl, 1234, pin=False)
d = bench.to_device(host, dev)
banks = bench.assemble_banks(d, wl)
fwd = model
if os.environ.get("PROFILE_GRAPHS"):
    from sam2_video_training_b200.graphs import GraphedMemoryAttention
    fwd = GraphedMemoryAttention(model)
for _ in range(3): bench.run_step(model, crit, opt, d, banks, wl, 1, fwd=fwd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): bench.run_step(model, crit, opt, d, banks, wl, 1, fwd=fwd)
e1.record(); torch.cuda.synchronize()
print("step wall (CUDA events, 3 steps): %.2f ms/step" % (e0.elapsed_time(e1) / 3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    bench.run_step(model, crit, opt, d, banks, wl, 1, fwd=fwd)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
