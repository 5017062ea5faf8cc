"""Same-box stock-PyTorch baseline (TEST / MEASUREMENT INFRASTRUCTURE ONLY -- lives under oracle/, never imported
by the product package).

SURVEY.md section 8d asks for "the same reference modules on the B200 (.cuda(), fp32 and bf16-autocast, torch SDPA)"
as the number the hand-written kernels must beat.  /root/reference does not exist on the GPU box, so this file
restates the reference's COMPOSITION with the stock PyTorch operators it calls -- ``F.linear``, ``F.layer_norm``,
``F.scaled_dot_product_attention`` (transformer.py:304-306), complex-multiply RoPE (position_encoding.py:204-239),
``memory + pos`` re-added in every layer (memory_attention.py:66-81) -- on the reference's state_dict keys, and runs
the bench.py step (same synthetic inputs, same loss weights, fused AdamW) with it.  The loss leg is the oracle's torch
restatement of losses.py:20-248 executed on the GPU (a few dozen ATen kernels per frame, like the reference's).

``python oracle/torch_gpu_baseline.py [workload] [--steps K]`` prints one line per precision mode.
``tests/test_oracle_golden.py::test_torch_baseline_matches_oracle`` pins this composition to the oracle on the CPU.
"""
from __future__ import annotations

import math
import os
import sys
from typing import Dict

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

Tensor = torch.Tensor


def rope_table_complex(n_tokens: int, dim: int = 256, theta: float = 10000.0, device=None) -> Tensor:
    """position_encoding.py:185-201: complex table [N, dim/2] = cat(polar(1, x f), polar(1, y f))."""
    w = int(round(math.sqrt(n_tokens)))
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4, device=device)[: dim // 4].float() / dim))
    t = torch.arange(n_tokens, dtype=torch.float32, device=device)
    ang = torch.cat([torch.outer(t % w, freqs), torch.outer(torch.div(t, w, rounding_mode="floor"), freqs)], dim=-1)
    return torch.polar(torch.ones_like(ang), ang)


def rotate(x: Tensor, table: Tensor) -> Tensor:
    """position_encoding.py:204-239: fp32 complex view, multiply, cast back; keys tile the table."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    r = xc.shape[-2] // table.shape[0]
    tb = table.repeat(r, 1) if r > 1 else table
    return torch.view_as_real(xc * tb).flatten(-2).type_as(x)


def rope_attention(p: Dict[str, Tensor], pre: str, q: Tensor, k: Tensor, v: Tensor, n_exclude: int, table: Tensor) -> Tensor:
    """transformer.py:275-311, one head of 256."""
    q = F.linear(q, p[pre + "q_proj.weight"], p[pre + "q_proj.bias"]).unsqueeze(1)
    k = F.linear(k, p[pre + "k_proj.weight"], p[pre + "k_proj.bias"]).unsqueeze(1)
    v = F.linear(v, p[pre + "v_proj.weight"], p[pre + "v_proj.bias"]).unsqueeze(1)
    n_rope = k.shape[-2] - n_exclude
    q = rotate(q, table)
    if n_rope > 0:
        k = torch.cat([rotate(k[:, :, :n_rope], table), k[:, :, n_rope:]], dim=-2)
    o = F.scaled_dot_product_attention(q, k, v).squeeze(1)
    return F.linear(o, p[pre + "out_proj.weight"], p[pre + "out_proj.bias"])


def memory_attention(p: Dict[str, Tensor], curr: Tensor, memory: Tensor, curr_pos: Tensor, memory_pos: Tensor,
                     num_obj_ptr_tokens: int, table: Tensor) -> Tensor:
    """memory_attention.py:58-99,119-169 with the shipped yaml flags (dropout off)."""
    x = (curr + 0.1 * curr_pos).transpose(0, 1)
    mem, mpos = memory.transpose(0, 1), memory_pos.transpose(0, 1)
    for i in range(4):
        pre = f"layers.{i}."
        t2 = F.layer_norm(x, (256,), p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        x = x + rope_attention(p, pre + "self_attn.", t2, t2, t2, 0, table)
        t2 = F.layer_norm(x, (256,), p[pre + "norm2.weight"], p[pre + "norm2.bias"])
        x = x + rope_attention(p, pre + "cross_attn_image.", t2, mem + mpos, mem, num_obj_ptr_tokens, table)
        t2 = F.layer_norm(x, (256,), p[pre + "norm3.weight"], p[pre + "norm3.bias"])
        h = F.relu(F.linear(t2, p[pre + "linear1.weight"], p[pre + "linear1.bias"]))
        x = x + F.linear(h, p[pre + "linear2.weight"], p[pre + "linear2.bias"])
    return F.layer_norm(x, (256,), p["norm.weight"], p["norm.bias"]).transpose(0, 1)


def run_step(p, opt, d, banks, wl, table, autocast: bool):
    from oracle import losses_oracle as lo
    import bench
    T = wl["T"]
    for t in range(1, T):
        mem, pos, n_ptr = banks[t - 1]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = memory_attention(p, d["curr"][t - 1], mem, d["curr_pos"], pos, n_ptr, table)
        out.backward(d["grad_out"][t - 1].to(out.dtype))
        pos.grad = None
    for ci in range(wl["clips"]):
        xs = [d["logits"][ci * T + f].requires_grad_(True) for f in range(T)]
        ip = d["iou"][ci].requires_grad_(True)
        l = lo.multistep_loss(xs, d["targets"][ci], [ip[f] for f in range(T)], dict(bench.LOSS_W), iou_use_l1_loss=True)
        l["total_loss"].backward()
        for x in xs:
            x.grad = None
        ip.grad = None
    opt.step()
    opt.zero_grad(set_to_none=False)


def main():
    import argparse
    import json
    import bench
    from oracle import attention_oracle as ao
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="cfg2_endovis18_384px_T10_7obj_x8clips")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--modes", default="bf16_autocast,fp32")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    d = bench.to_device(bench.make_host_inputs(wl, 1234, pin=False), dev)
    banks = bench.assemble_banks(d, wl)
    table = rope_table_complex(wl["grid"] ** 2, device=dev)
    for mode in args.modes.split(","):
        p = {k: v.to(dev).requires_grad_(True) for k, v in ao.init_params(seed=0).items()}
        opt = torch.optim.AdamW(list(p.values()), lr=1e-5, fused=True)
        for _ in range(args.warmup):
            run_step(p, opt, d, banks, wl, table, mode == "bf16_autocast")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run_step(p, opt, d, banks, wl, table, mode == "bf16_autocast")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"impl": "stock_pytorch_same_box", "mode": mode, "workload": args.workload, "ms_per_step": ms,
                          "clip_frames_per_s": wl["clips"] * wl["T"] / (ms * 1e-3), "steps": args.steps,
                          "tflops_algorithmic": bench.algorithmic_flops(wl) / (ms * 1e-3) / 1e12,
                          "sdpa": "F.scaled_dot_product_attention default backend selection", "torch": torch.__version__,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
        del p, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
