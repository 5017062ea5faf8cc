"""One call of each sam2b200_gemm(_ex) configuration of a cfg2 layer (R = 56 x 576 rows, memory bank of 4060 tokens), for
`ncu --set full -k regex:gemm_kernel` captures.  Inputs are re-created between the calls so that nothing is L2-resident by accident."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import fused_stack as fs
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0"); BF16 = torch.bfloat16
b, n, m = 56, 576, 4060
r = b * n
g = torch.Generator(device="cuda").manual_seed(1)
table = compute_axial_cis(dim=256, end_x=24, end_y=24).to(dev)
def rnd(*s, scale=1.0): return (torch.randn(*s, device=dev, generator=g) * scale).to(BF16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for it in range(2):
    y, h, dqkv, mem = rnd(r, 256), rnd(r, 2048), rnd(r, 768), rnd(b * m, 64)
    o32 = torch.randn(r, 64, device=dev, generator=g)
    wqkv, w1, w2, wo, weff, wk = rnd(768, 256, scale=1 / 16), rnd(2048, 256, scale=1 / 16), rnd(256, 2048, scale=1 / 45), rnd(256, 256, scale=1 / 16), rnd(256, 64, scale=1 / 16), rnd(256, 64, scale=1 / 8)
    b768, b2048, b256 = (torch.zeros(k, device=dev) for k in (768, 2048, 256))
    calls = [lambda: fs.gemm_ex(y, wqkv, 3, 256, bias=b768, table=table, rope_outs=2, rows_per_item=n, n_rope_rows=n),      # q|k|v head + RoPE   <128,0,1,2>
             lambda: fs.gemm_ex(y, w1, 1, 2048, bias=b2048, relu=True),                                                      # linear1 + ReLU      <256,0,1,1>
             lambda: fs.gemm(h, w2, bias=b256),                                                                              # linear2             <256,0,2,1>
             lambda: fs.gemm(h, w1, nn=True),                                                                                # d linear1           <256,1,2,1>
             lambda: fs.gemm(dqkv, wqkv, nn=True),                                                                           # d q|k|v (K = 768)   <256,1,2,1>
             lambda: fs.gemm(y, wo, nn=True),                                                                                # d out_proj          <256,1,1,1>
             lambda: fs.gemm(y, weff, nn=True, dot_rows=o32),                                                                # d folded + Delta    <64,1,1,1>
             lambda: fs.gemm(mem, wk, bias=b256, table=table, rows_per_item=m, n_rope_rows=(m // n) * n)]                    # memory keys + RoPE  <128,0,1,2>
    for c in calls:
        flush.zero_()
        c()
    torch.cuda.synchronize()
print("done")
