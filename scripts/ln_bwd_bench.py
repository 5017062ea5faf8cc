"""ln_bwd kernel at the cfg2 row count (B N = 32256 rows x 256), CUDA events; SAM2B200_LN_BWD_BLOCKS_PER_SM selects the grid."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32256
g = torch.Generator(device="cuda").manual_seed(0)
nset = 6   # rotate inputs: 6 x 132 MB > L2
sets = []
for _ in range(nset):
    x = torch.randn(rows, 256, device=dev, generator=g)
    dy = torch.randn(rows, 256, device=dev, generator=g).to(torch.bfloat16)
    gin = torch.randn(rows, 256, device=dev, generator=g)
    mean, rstd = x.mean(-1), 1 / (x.var(-1, unbiased=False) + 1e-5).sqrt()
    sets.append((dy, x, mean.contiguous(), rstd.contiguous(), gin))
gamma = torch.ones(256, device=dev)
dgam, dbet, dbias = (torch.zeros(256, device=dev) for _ in range(3))
def run(k):
    dy, x, mean, rstd, gin = sets[k % nset]
    return fs.ln_bwd(dy, x, mean, rstd, gamma, gin, dgam, dbet, dbias=dbias)
for k in range(6): run(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
it = 60
e0.record()
for k in range(it): run(k)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / it * 1e3
byt = rows * 256 * (2 + 4 + 4 + 4 + 2)
print(f"blocks/SM={os.environ.get('SAM2B200_LN_BWD_BLOCKS_PER_SM','2')} rows={rows}: {us:.1f} us per ln_bwd (+ partial reduce), {byt/us/1e3:.0f} GB/s algorithmic")
