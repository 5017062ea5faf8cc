"""Memory encoder forward + backward per frame (cfg2: 56 objects, 384 px masks -> 24 x 24 x 64 memory features; cfg4: 4 objects,
1024 px): this repo's MemoryEncoder vs the unmodified reference class from baseline/_ref on the same GPU (fp32 and bf16 autocast).
CUDA events, 10 iterations after 3 warm-ups."""
import os, sys, types, importlib.util, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sam2_video_training_b200.modeling.memory_encoder import build_memory_encoder
dev = torch.device("cuda:0")
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
ref_cls = None
if bench.load_reference() is not None:
    mdl = os.path.join(bench.REF_ROOT, "sam2_video", "model", "modeling")
    spec = importlib.util.spec_from_file_location("sam2.modeling.memory_encoder", os.path.join(mdl, "memory_encoder.py"))
    me = importlib.util.module_from_spec(spec); sys.modules["sam2.modeling.memory_encoder"] = me; spec.loader.exec_module(me)
    pe = sys.modules["sam2.modeling.position_encoding"]
    def build_ref():
        return me.MemoryEncoder(out_dim=64, position_encoding=pe.PositionEmbeddingSine(num_pos_feats=64, normalize=True, scale=None, temperature=10000),
                                mask_downsampler=me.MaskDownSampler(kernel_size=3, stride=2, padding=1),
                                fuser=me.Fuser(layer=me.CXBlock(dim=256, kernel_size=7, padding=3, layer_scale_init_value=1e-6, use_dwconv=True), num_layers=2))
    ref_cls = build_ref
for name, b, grid in (("cfg2 (56 objects, 384 px)", 56, 24), ("cfg3 (13 objects, 512 px)", 13, 32), ("cfg4 (4 objects, 1024 px)", 4, 64)):
    g = torch.Generator(device="cuda").manual_seed(0)
    pix = torch.randn(b, 256, grid, grid, device=dev, generator=g)
    masks = torch.randn(b, 1, 16 * grid, 16 * grid, device=dev, generator=g) * 4
    gout = torch.randn(b, 64, grid, grid, device=dev, generator=g)
    torch.manual_seed(0)
    mine = build_memory_encoder().to(dev).train()
    def run(model, autocast=False):
        model.zero_grad(set_to_none=True)
        p = pix.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = model(p, torch.sigmoid(masks) * 20 - 10, skip_mask_sigmoid=True)["vision_features"]
        out.backward(gout.to(out.dtype))
    t_mine = timeit(lambda: run(mine))
    line = f"{name}: this repo {t_mine:7.3f} ms fwd+bwd"
    if ref_cls is not None:
        torch.manual_seed(0)
        ref = ref_cls().to(dev).train()
        line += f" | reference class on this GPU: fp32 {timeit(lambda: run(ref)):7.3f} ms, bf16 autocast {timeit(lambda: run(ref, True)):7.3f} ms"
    print(line, flush=True)
