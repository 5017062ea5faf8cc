"""Probe (GPU box): which in-place / out= forms of bf16 x bf16 -> fp32 GEMMs does this torch build support?"""
import torch
dev = torch.device("cuda:0")
a = torch.randn(512, 256, device=dev).bfloat16(); b = torch.randn(512, 384, device=dev).bfloat16()
ref = a.float().t() @ b.float()
flat = torch.zeros(256 * 384 + 64, device=dev)
view = flat[64:].view(256, 384)
def attempt(name, fn):
    try:
        flat.zero_(); view.fill_(1.0)
        r = fn()
        torch.cuda.synchronize()
        tgt = view if r is None or r.data_ptr() == view.data_ptr() else r
        err = float((tgt - 1.0 - ref).abs().max() / ref.abs().max())
        print(f"{name}: ok, same_storage={r is not None and r.data_ptr() == view.data_ptr()}, rel err vs (1 + a^T b) = {err:.2e}")
    except Exception as e:
        print(f"{name}: FAILED {type(e).__name__}: {str(e)[:200]}")
attempt("addmm(view, a.t, b, out_dtype=f32, out=view)", lambda: torch.addmm(view, a.t(), b, out_dtype=torch.float32, out=view))
attempt("addmm(view, a.t, b, out_dtype=f32)", lambda: torch.addmm(view, a.t(), b, out_dtype=torch.float32))
attempt("view.addmm_(a.t, b)", lambda: view.addmm_(a.t(), b))
attempt("mm(a.t, b, out_dtype=f32, out=view) (overwrites)", lambda: torch.mm(a.t(), b, out_dtype=torch.float32, out=view))
