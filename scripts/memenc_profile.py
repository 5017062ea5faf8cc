"""torch-profiler kernel table of one memory-encoder forward + backward at the cfg2 frame shape (56 objects, 384 px)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200.modeling.memory_encoder import build_memory_encoder
dev = torch.device("cuda:0")
b, grid = 56, 24
g = torch.Generator(device="cuda").manual_seed(0)
pix = torch.randn(b, 256, grid, grid, device=dev, generator=g)
masks = torch.randn(b, 1, 16 * grid, 16 * grid, device=dev, generator=g) * 4
gout = torch.randn(b, 64, grid, grid, device=dev, generator=g)
torch.manual_seed(0)
m = build_memory_encoder().to(dev).train()
def run():
    m.zero_grad(set_to_none=True)
    p = pix.clone().requires_grad_(True)
    m(p, torch.sigmoid(masks) * 20 - 10, skip_mask_sigmoid=True)["vision_features"].backward(gout)
for _ in range(3): run()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=90))
