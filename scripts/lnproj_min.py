import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_video_training_b200 import fused_stack as fs
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
dev = torch.device("cuda:0")
r = int(sys.argv[1]); rope = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 576 if r % 576 == 0 else r
x = torch.randn(r, 256, device=dev); res = torch.randn(r, 256, device=dev).to(torch.bfloat16)
g, b = torch.ones(256, device=dev), torch.zeros(256, device=dev)
table = compute_axial_cis(dim=256, end_x=24, end_y=24).to(dev)
print("table", table.dtype, tuple(table.shape), hex(table.data_ptr()), table.is_contiguous(), flush=True)
for nout, n_out, width, relu, ropes in ((256, 1, 256, False, 1), (768, 3, 256, False, 2), (2048, 1, 2048, True, 0)):
    w = (torch.randn(nout, 256, device=dev) / 16).to(torch.bfloat16); bias = torch.zeros(nout, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        outs, y, xn, mean, rstd = fs.ln_proj(x, res, g, b, w, bias, n_out, out_width=width, relu=relu, table=table if (rope and ropes) else None,
                                             rope_outs=ropes if rope else 0, rows_per_item=576, n_rope_rows=576)
    torch.cuda.synchronize()
    if not rope:
        ref = y.float() @ w.float().t()
        if relu: ref = torch.relu(ref)
        got = torch.cat([o.float() for o in outs], 1)
        print(r, rope, nout, "ok", float((got - ref).norm() / ref.norm()), flush=True)
    else:
        print(r, rope, nout, "ran", flush=True)
