"""Micro-benchmark of the memory-bank assembly (GPU box): sam2b200_bank_gather through memory_bank.assemble_memory vs
the same assembly written with the reference's torch ops (flatten/permute/add/cat, sam2_base.py:597-692) on the GPU.
Algorithmic bytes: read 2 x S frames x B x 64 x HW x 4 B + pointers, write 2 x M x B x 64 x 4 B.
usage: python scripts/bank_bench.py [B,grid,frames,pointers ...]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sam2_video_training_b200 import memory_bank as mb


def torch_assembly(cfg, frames, pointers, tpos, proj, num_frames):
    mem, pos = [], []
    for t_pos, o in frames:
        mem.append(o["maskmem_features"].flatten(2).permute(2, 0, 1))
        pos.append(o["maskmem_pos_enc"][-1].flatten(2).permute(2, 0, 1) + tpos[cfg.num_maskmem - t_pos - 1])
    pos_list, ptrs = zip(*pointers)
    obj_ptrs = torch.stack(ptrs, dim=0)
    b = obj_ptrs.shape[1]
    obj_pos = proj(mb.get_1d_sine_pe(torch.tensor(pos_list, dtype=torch.float32, device=obj_ptrs.device) / (min(num_frames, 16) - 1), 256))
    obj_pos = obj_pos.unsqueeze(1).expand(-1, b, 64)
    obj_ptrs = obj_ptrs.reshape(-1, b, 4, 64).permute(0, 2, 1, 3).flatten(0, 1)
    obj_pos = obj_pos.repeat_interleave(4, dim=0)
    return torch.cat(mem + [obj_ptrs], 0), torch.cat(pos + [obj_pos], 0)


def bench(b, grid, nf, nptr, iters=20):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    cfg = mb.BankConfig()
    num_frames = 32
    frame_idx = 20
    od = {"cond_frame_outputs": {0: None}, "non_cond_frame_outputs": {}}

    def frame():
        return {"maskmem_features": torch.randn(b, 64, grid, grid, device=dev, generator=g),
                "maskmem_pos_enc": [torch.randn(b, 64, grid, grid, device=dev, generator=g)],
                "obj_ptr": torch.randn(b, 256, device=dev, generator=g)}
    od["cond_frame_outputs"][0] = frame()
    for t in range(frame_idx - max(nf - 1, nptr - 1), frame_idx):
        od["non_cond_frame_outputs"][t] = frame()
    tpos = torch.randn(7, 1, 1, 64, device=dev, generator=g)
    proj = torch.nn.Linear(256, 64).to(dev)
    frames, pointers = mb.select_bank_entries(cfg, frame_idx, od, num_frames, True)
    frames, pointers = frames[:nf] if len(frames) > nf else frames, pointers[:nptr]

    def ours():
        return mb.assemble_memory(cfg, frame_idx, od, num_frames, tpos, proj, training=True)

    def theirs():
        return torch_assembly(cfg, frames, pointers, tpos, proj, num_frames)

    m0, p0, n0 = ours()
    m1, p1 = theirs()
    assert torch.equal(m0, m1) and torch.allclose(p0, p1, atol=2e-6, rtol=0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn):
        ts = []
        for _ in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts[3:])[len(ts[3:]) // 2]

    t_ours, t_torch = timeit(ours), timeit(theirs)
    # the gather kernels alone (device time, back-to-back launches of the autograd function with prepared arguments)
    feats = [o["maskmem_features"] for _, o in frames]; pos = [o["maskmem_pos_enc"][-1] for _, o in frames]
    ptrs = [p_ for _, p_ in pointers]
    tp = tpos.reshape(7, 64)[:len(frames)].contiguous(); op = torch.randn(len(ptrs), 64, device=dev)
    for _ in range(3): mb._BankGatherFn.apply(len(frames), len(ptrs), 256, tp, op, *feats, *pos, *ptrs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): mb._BankGatherFn.apply(len(frames), len(ptrs), 256, tp, op, *feats, *pos, *ptrs)
    e1.record(); torch.cuda.synchronize()
    t_kernel = e0.elapsed_time(e1) / 20
    s = len(frames)
    byts = 2 * s * b * 64 * grid * grid * 4 + len(pointers) * b * 256 * 4 + 2 * m0.numel() * 4
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(f"B={b} grid={grid} frames={s} pointers={len(pointers)} M={m0.shape[0]}: bank_gather path {t_ours*1e3:7.1f} us "
          f"({byts/t_ours/1e6:5.0f} GB/s, {byts/t_ours/1e6/pk:5.1%} of {pk:.0f}; includes the tiny torch ops for the pointer positions) | "
          f"torch ops {t_torch*1e3:7.1f} us | x{t_torch/t_ours:.1f} | gather kernels alone {t_kernel*1e3:.1f} us = {byts/t_kernel/1e6:.0f} GB/s "
          f"({byts/t_kernel/1e6/pk:.1%})", flush=True)


if __name__ == "__main__":
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(56, 24, 7, 16), (13, 32, 7, 16), (4, 64, 7, 16), (1, 24, 7, 10)]
    for s in shapes:
        bench(*s)
