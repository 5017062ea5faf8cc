"""Probe (GPU box): which sequence makes a later CUDA-graph capture of the direct-gradient path fail?"""
import os, sys, subprocess, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def scenario(name):
    from sam2_video_training_b200 import ddp
    from sam2_video_training_b200.graphs import GraphedMemoryAttention
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    dev = torch.device("cuda:0")
    model = build_memory_attention(dropout=0.0).to(dev).train()
    ddp.attach_grad_bucket(model)
    gm = GraphedMemoryAttention(model)
    n, b = 64, 2
    def call(m, pos_grad, keep=False):
        curr = torch.randn(n, b, 256, device=dev); mem = torch.randn(m, b, 64, device=dev)
        cpos = torch.randn(n, b, 256, device=dev); mpos = torch.randn(m, b, 64, device=dev).requires_grad_(pos_grad)
        o = gm(curr, mem, cpos, mpos, 8)
        o.backward(torch.randn_like(o))
        return o if keep else None
    if name == "nograd_first":
        call(136, False)
    elif name == "grad_then_nograd":
        call(136, True); call(136, False)
    elif name == "grad_then_grad_other_shape":
        call(136, True); call(200, True)
    elif name == "grad_keepalive_then_grad_other_shape":
        keep = call(136, True, keep=True); call(200, True)
    elif name == "grad_update_then_grad_other_shape":
        call(136, True)
        with torch.no_grad():
            for p in model.parameters(): p.add_(0.01)
        call(136, True); call(200, True)
    elif name == "nograd_then_nograd_other_shape":
        call(136, False); call(200, False)
    torch.cuda.synchronize()
    print("OK")

if __name__ == "__main__":
    if len(sys.argv) > 1:
        scenario(sys.argv[1])
    else:
        for s in ["nograd_first", "grad_then_nograd", "grad_then_grad_other_shape", "grad_keepalive_then_grad_other_shape",
                  "grad_update_then_grad_other_shape", "nograd_then_nograd_other_shape"]:
            r = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True)
            err = [l for l in r.stderr.splitlines() if "Error" in l]
            print(f"{s}: {'OK' if r.returncode == 0 else 'FAIL ' + (err[0][:120] if err else '')}", flush=True)
