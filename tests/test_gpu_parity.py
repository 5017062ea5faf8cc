"""-m gpu parity tests: the CUDA path (through the C ABI) against the CPU oracle and against the
golden vectors generated from the unmodified reference.  Tolerances are the ones BASELINE.json
states: attention outputs <= 1e-2 relative (bf16 vs fp32 reference), losses <= 1e-3 relative,
gradient cosine similarity >= 0.999."""
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import attention_oracle as ao
from oracle import detgen
from oracle import losses_oracle as lo

pytestmark = pytest.mark.gpu

W_FOCAL = {"loss_mask": 20, "loss_dice": 1, "loss_iou": 1, "loss_class": 0}
ATTN_REL_TOL = 1e-2     # relative L2 error of attention outputs (north star)
LOSS_REL_TOL = 1e-3     # relative error of loss values (north star)
GRAD_COS_TOL = 0.999    # gradient cosine similarity (north star)


def rel_l2(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def cosine(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp(min=1e-30))


@pytest.fixture(scope="module")
def dev():
    from sam2_video_training_b200 import _lib
    lib = _lib.load()
    assert lib.sam2b200_check_device(0) == 0, lib.sam2b200_last_error()
    return torch.device("cuda:0")


def test_selftest_binary():
    """The torch-free self-test of the C ABI (csrc/selftest.cu) against its own fp64 CPU restatement."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sam2_video_training_b200",
                       "sam2b200_selftest")
    if not os.path.exists(exe):
        pytest.skip("selftest binary not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]


# ------------------------------------------------------------------ attention core
@pytest.mark.parametrize("b,grid,nf,nptr,nsplit", [
    (1, 8, 1, 0, 1),        # N = M = 64: one tile
    (2, 8, 3, 12, 1),       # ragged M = 204
    (1, 24, 1, 4, 0),       # cfg1 frame 1: N = 576, M = 580 (library split heuristic)
    (1, 24, 7, 28, 0),      # cfg1 steady state: M = 4060
    (1, 24, 7, 28, 1),      # same without split-KV
    (3, 16, 2, 8, 2),       # forced split
    (2, 32, 1, 0, 1),       # self-attention at 512 px: N = M = 1024
])
def test_rope_attention_core(dev, b, grid, nf, nptr, nsplit):
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    from sam2_video_training_b200.ops import RopeAttentionFn
    n = grid * grid
    m = nf * n + nptr
    g = torch.Generator().manual_seed(1234 + n + m)
    q = torch.randn(b, n, 256, generator=g) * 1.5
    k = torch.randn(b, m, 256, generator=g) * 1.5
    v = torch.randn(b, m, 256, generator=g)
    do = torch.randn(b, n, 256, generator=g)
    # the kernel sees bf16 operands: give the oracle the same rounded values (error budget = kernel only)
    qb, kb, vb, dob = (t.to(torch.bfloat16) for t in (q, k, v, do))
    qo, ko, vo = (t.double().requires_grad_(True) for t in (qb, kb, vb))
    ref = ao.core_attention(qo, ko, vo, nptr)
    ref.backward(dob.double())
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (qb, kb, vb))
    out = RopeAttentionFn.apply(qd, kd, vd, table, nptr, nsplit)
    out.backward(dob.to(dev))
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16
    assert rel_l2(out, ref) < ATTN_REL_TOL
    for name, mine, theirs in (("dq", qd.grad, qo.grad), ("dk", kd.grad, ko.grad), ("dv", vd.grad, vo.grad)):
        assert torch.isfinite(mine).all(), name
        assert cosine(mine, theirs) > GRAD_COS_TOL, (name, cosine(mine, theirs))
        assert rel_l2(mine, theirs) < 3e-2, (name, rel_l2(mine, theirs))


def test_rope_only(dev):
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    from sam2_video_training_b200.ops import rope_apply
    grid, b, nf, p = 8, 2, 2, 6
    n = grid * grid
    x = torch.randn(b, nf * n + p, 256)
    cos, sin = ao.axial_rope_table(n)
    ref = torch.cat([ao.apply_axial_rope(x[:, :nf * n], cos, sin), x[:, nf * n:]], dim=1)
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    got = rope_apply(x.to(dev), table, nf * n, out_dtype=torch.float32)
    assert torch.allclose(got.cpu(), ref, atol=2e-6, rtol=1e-6)
    back = rope_apply(got, table, nf * n, inverse=True, out_dtype=torch.float32)
    assert torch.allclose(back.cpu(), x, atol=5e-6, rtol=1e-5)  # conjugate rotation is the inverse
    gotb = rope_apply(x.to(dev).to(torch.bfloat16), table, nf * n)
    assert rel_l2(gotb, ref) < 6e-3


# ------------------------------------------------------------------ full MemoryAttention stack
def _load_params(model, params):
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(params[n])


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("tag", ["g4_b2_f2_p8", "g8_b3_f3_p12", "g12_b1_f1_p0"])
def test_memory_attention_vs_reference_golden(dev, golden_dir, tag, fused):
    """Against outputs of the UNMODIFIED reference (fp32 CPU) stored in tests/golden.

    The golden inputs/weights are smooth analytic functions (sin of the flat index): keys / values are
    highly correlated and many ReLU pre-activations sit near 0, so the *gradients* are ill-conditioned
    with respect to bf16 rounding -- rounding only the operands of the reference's own nn.Linear calls
    to bf16 (attention core left in fp32) already drops individual parameter-gradient cosines to
    0.92-0.99 (measured with the oracle; see DESIGN.md "Numerics").  The forward output keeps the
    north-star tolerance; gradients get the strict test on random inputs below."""
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    g = np.load(os.path.join(golden_dir, f"attn_{tag}.npz"))
    grid, batch, nf, nptr = int(g["grid"]), int(g["batch"]), int(g["n_frames"]), int(g["n_ptr"])
    model = build_memory_attention().to(dev).eval()
    model.use_fused_stack = fused
    _load_params(model, detgen.det_params(detgen.param_shapes()))
    inp = {k: v.to(dev) for k, v in detgen.attention_inputs(grid, batch, nf, nptr).items()}
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("curr", "curr_pos", "memory", "memory_pos")}
    out = model(curr=[leaves["curr"]], curr_pos=[leaves["curr_pos"]], memory=leaves["memory"],
                memory_pos=leaves["memory_pos"], num_obj_ptr_tokens=nptr)
    out.backward(inp["grad_out"])
    torch.cuda.synchronize()
    assert out.shape == (grid * grid, batch, 256) and out.dtype == torch.float32
    assert rel_l2(out, torch.from_numpy(g["out"])) < ATTN_REL_TOL
    for k in ("curr", "curr_pos", "memory", "memory_pos"):
        assert cosine(leaves[k].grad, torch.from_numpy(g["d_" + k])) > 0.995, k
    # parameter gradients: asserted at the north-star 0.999 on the WELL-CONDITIONED reference fixtures below
    # (test_memory_attention_vs_reference_refinit_golden); on these analytic ones only finiteness
    for _, p in model.named_parameters():
        assert torch.isfinite(p.grad).all()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("tag", ["g24_b1_f7_p28", "g8_b3_f3_p12"])
def test_memory_attention_vs_reference_refinit_golden(dev, golden_dir, tag, fused):
    """North-star tolerances against the UNMODIFIED REFERENCE itself (not the oracle): the reference stack with its own
    random init (torch.manual_seed(0)) on N(0,1) inputs, BASELINE.json configs[0] steady-state shape (24x24 tokens, 7
    memory frames + 28 pointer tokens) and a ragged 8x8 / 3-object case -- oracle/make_golden.py::
    golden_attention_refinit.  Weights and inputs are regenerated and proven identical by checksum."""
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    g = np.load(os.path.join(golden_dir, f"attn_refinit_{tag}.npz"))
    grid, batch, nf, nptr, seed = (int(g[k]) for k in ("grid", "batch", "n_frames", "n_ptr", "seed"))
    params = ao.reference_init_params(0)
    for n, s_ in zip([str(n) for n in g["param_names"]], g["weight_abs_sums"]):
        assert abs(float(params[n].double().abs().sum()) - s_) <= 1e-9 * max(s_, 1.0), n
    inp = ao.random_inputs(grid, batch, nf, nptr, seed)
    for k, s_ in zip(("curr", "curr_pos", "memory", "memory_pos", "grad_out"), g["input_abs_sums"]):
        assert abs(float(inp[k].double().abs().sum()) - s_) <= 1e-9 * s_, k
    model = build_memory_attention().to(dev).eval()
    model.use_fused_stack = fused
    _load_params(model, params)
    leaves = {k: inp[k].to(dev).requires_grad_(True) for k in ("curr", "curr_pos", "memory", "memory_pos")}
    out = model(curr=[leaves["curr"]], curr_pos=[leaves["curr_pos"]], memory=leaves["memory"],
                memory_pos=leaves["memory_pos"], num_obj_ptr_tokens=nptr)
    out.backward(inp["grad_out"].to(dev))
    torch.cuda.synchronize()
    assert rel_l2(out, torch.from_numpy(g["out"])) < ATTN_REL_TOL
    for k in ("curr", "memory", "memory_pos"):
        assert cosine(leaves[k].grad, torch.from_numpy(g["d_" + k])) > GRAD_COS_TOL, (k, cosine(leaves[k].grad, torch.from_numpy(g["d_" + k])))
    assert cosine(leaves["curr_pos"].grad, torch.from_numpy(0.1 * g["d_curr"])) > GRAD_COS_TOL
    named = dict(model.named_parameters())
    for key in g.files:
        if key.startswith("dparam:"):
            c = cosine(named[key[7:]].grad, torch.from_numpy(g[key]))
            assert c > GRAD_COS_TOL, (key, c)
    for n, s_ in zip([str(n) for n in g["param_names"]], g["param_grad_abs_sums"]):
        mine = float(named[n].grad.double().abs().sum())
        assert abs(mine - s_) <= 5e-2 * max(abs(s_), 1e-3), (n, mine, s_)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("grid,b,nf,nptr", [(24, 1, 7, 28), (16, 2, 3, 12)])
def test_memory_attention_random_vs_oracle(dev, grid, b, nf, nptr, fused):
    """BASELINE.json configs[0] shape (24x24 tokens, 1 object, steady-state bank of 7 frames + 7 pointers)
    and a smaller ragged case; random-init weights, random inputs, against the fp32 oracle.
    North-star tolerances: output <= 1e-2 relative, gradient cosine >= 0.999 (inputs, and all parameter
    gradients taken together); single parameter tensors behind the ReLU (linear1 / norm3) reach
    ~0.9988 because bf16 rounding flips ReLU units whose pre-activation is ~0 -- any bf16 path does."""
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    params = ao.init_params(seed=0)
    n, m = grid * grid, nf * grid * grid + nptr
    g = torch.Generator().manual_seed(7)
    inputs = dict(curr=torch.randn(n, b, 256, generator=g), curr_pos=torch.randn(n, b, 256, generator=g) * 0.7,
                  memory=torch.randn(m, b, 64, generator=g), memory_pos=torch.randn(m, b, 64, generator=g) * 0.7)
    gout = torch.randn(n, b, 256, generator=g)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    lo_ = {k: v.clone().requires_grad_(True) for k, v in inputs.items()}
    ref = ao.memory_attention(po, lo_["curr"], lo_["memory"], lo_["curr_pos"], lo_["memory_pos"], nptr)
    ref.backward(gout)
    model = build_memory_attention().to(dev).eval()
    model.use_fused_stack = fused
    _load_params(model, params)
    ld = {k: v.to(dev).clone().requires_grad_(True) for k, v in inputs.items()}
    out = model(ld["curr"], ld["memory"], ld["curr_pos"], ld["memory_pos"], nptr)
    out.backward(gout.to(dev))
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < ATTN_REL_TOL
    for k in ld:
        assert cosine(ld[k].grad, lo_[k].grad) > GRAD_COS_TOL, k
    names = [nm for nm, _ in model.named_parameters()]
    mine = torch.cat([p.grad.flatten().cpu() for _, p in model.named_parameters()])
    theirs = torch.cat([po[nm].grad.flatten() for nm in names])
    assert cosine(mine, theirs) > GRAD_COS_TOL
    worst = min((cosine(p.grad, po[nm].grad), nm) for nm, p in model.named_parameters())
    assert worst[0] > 0.998, worst


def test_raw_memory_cross_attention_dropout_matches_explicit_mask(dev):
    """Attention-probability dropout on the raw-memory path: rows of the dropped matrix do not sum to 1, so the forward
    returns their sums (factor of the value bias) and the backward takes the per-query constant dO . bv.  Against fp32
    torch of the reference formulation softmax -> mask / (1 - p) -> @ (mem Wv^T + bv) with the kernels' own mask."""
    from sam2_video_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(77)
    b, n, m, p_drop, site = 3, 144, 300, 0.1, 9
    seed = torch.tensor([0x0123456789ABCDE], dtype=torch.int64, device=dev)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    wv = (torch.randn(256, 64, device=dev, generator=g) / 8).to(torch.bfloat16)
    bv = torch.randn(256, device=dev, generator=g) * 0.5
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    keep = _keep_mask(dev, seed, site, p_drop, b * n * m).view(b, n, m)
    qf, kf, mf, wf, bf = (t.float().requires_grad_(True) for t in (q, k, mem, wv, bv))
    a = torch.softmax(qf @ kf.transpose(1, 2) / 16.0, dim=-1) * keep / (1 - p_drop)
    ref = a @ (mf @ wf.t() + bf)
    ref.backward(do.float())
    drop = (p_drop, seed, site)
    o64, o64_32, lse, rs = ops.attn_fwd_v64(q, k, mem, 1 / 16.0, drop=drop)
    assert rel_l2(rs, a.sum(-1)) < 1e-3
    out = o64_32 @ wv.float().t() + rs[..., None] * bv
    assert rel_l2(out, ref) < 5e-3
    do64 = (do.float() @ wv.float()).to(torch.bfloat16)
    c = do.float() @ bv
    delta = (do64.float() * o64_32).sum(-1) + c * rs
    dq, dk = ops.attn_bwd_v64(q, k, mem, do64, lse, delta.contiguous(), 1 / 16.0, grad_dtype=torch.float32, dp_bias=c.contiguous(), drop=drop)
    assert rel_l2(dq, qf.grad) < 1e-2, rel_l2(dq, qf.grad)
    assert rel_l2(dk, kf.grad) < 1e-2, rel_l2(dk, kf.grad)
    assert rel_l2(do.float().flatten(0, 1).t() @ o64_32.flatten(0, 1), wf.grad) < 5e-3
    assert rel_l2((rs[..., None] * do.float()).sum((0, 1)), bf.grad) < 1e-3


def test_fused_stack_raw_memory_cross_attention_vs_oracle(dev, monkeypatch):
    """Training-shaped call (memory detached, memory_pos trainable, no dropout, enough objects to fill the GPU): the fused
    stack runs its cross-attention on the raw 64-d memory features (attn_fwd_v64 / attn_bwd_v64, v_proj applied to the
    [B N, 64] result, no dV kernel).  Output and every gradient against the fp32 oracle of the reference formulation."""
    from sam2_video_training_b200 import fused_stack
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    calls = {"fwd": 0, "bwd": 0}
    f0, f1, b0 = fused_stack.attn_fwd_v64, fused_stack.attn_fwd_v64_proj, fused_stack.attn_bwd_v64
    monkeypatch.setattr(fused_stack, "attn_fwd_v64", lambda *a, **k: (calls.__setitem__("fwd", calls["fwd"] + 1), f0(*a, **k))[1])
    # (default: the variant with the folded output projection in its epilogue)
    monkeypatch.setattr(fused_stack, "attn_fwd_v64_proj", lambda *a, **k: (calls.__setitem__("fwd", calls["fwd"] + 1), f1(*a, **k))[1])
    monkeypatch.setattr(fused_stack, "attn_bwd_v64", lambda *a, **k: (calls.__setitem__("bwd", calls["bwd"] + 1), b0(*a, **k))[1])
    params = ao.init_params(seed=0)
    grid, b, nf, nptr = 8, 64, 5, 12        # 3 key blocks x 64 objects > #SMs: the persistent dK kernel
    n, m = grid * grid, nf * grid * grid + nptr
    g = torch.Generator().manual_seed(19)
    inputs = dict(curr=torch.randn(n, b, 256, generator=g), curr_pos=torch.randn(n, b, 256, generator=g) * 0.7,
                  memory=torch.randn(m, b, 64, generator=g), memory_pos=torch.randn(m, b, 64, generator=g) * 0.7)
    gout = torch.randn(n, b, 256, generator=g)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    lo_ = {k: v.clone().requires_grad_(k != "memory") for k, v in inputs.items()}
    ref = ao.memory_attention(po, lo_["curr"], lo_["memory"], lo_["curr_pos"], lo_["memory_pos"], nptr)
    ref.backward(gout)
    model = build_memory_attention().to(dev).eval()
    _load_params(model, params)
    ld = {k: v.to(dev).clone().requires_grad_(k != "memory") for k, v in inputs.items()}
    out = model(ld["curr"], ld["memory"], ld["curr_pos"], ld["memory_pos"], nptr)
    out.backward(gout.to(dev))
    torch.cuda.synchronize()
    if not fused_stack.NO_V64:                             # (SAM2B200_NO_V64=1 is the A/B switch back to the 256-d value path)
        assert calls == {"fwd": 4, "bwd": 8}, calls        # 4 layers: forward once, backward dK + dQ
    assert rel_l2(out, ref) < ATTN_REL_TOL
    for k in ("curr", "curr_pos", "memory_pos"):
        assert cosine(ld[k].grad, lo_[k].grad) > GRAD_COS_TOL, k
    names = [nm for nm, _ in model.named_parameters()]
    mine = torch.cat([p.grad.flatten().cpu() for _, p in model.named_parameters()])
    theirs = torch.cat([po[nm].grad.flatten() for nm in names])
    assert cosine(mine, theirs) > GRAD_COS_TOL
    for nm in ("layers.0.cross_attn_image.v_proj.weight", "layers.3.cross_attn_image.v_proj.bias",
               "layers.2.cross_attn_image.k_proj.weight", "layers.1.cross_attn_image.q_proj.bias"):
        p = dict(model.named_parameters())[nm]
        assert cosine(p.grad, po[nm].grad) > GRAD_COS_TOL, nm


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_inputs_as_under_mixed_precision_training(dev, dtype):
    """configs/best.yaml:103 trains with 16-bit mixed precision: the tracker then hands MemoryAttention half-precision features.  The
    drop-in takes them (no fp32 copy is required from the caller), returns the fp32 output the reference's autocast LayerNorm returns,
    and gradients of the inputs' own dtype -- identical to the run on the same values given as fp32."""
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    torch.manual_seed(4)
    model = build_memory_attention(dropout=0.0).to(dev).train()
    g = torch.Generator(device="cuda").manual_seed(9)
    n, b, p = 64, 3, 8
    m = 2 * n + p
    mk = lambda *s: torch.randn(*s, device=dev, generator=g).to(dtype)
    curr, cpos, mem, mpos = mk(n, b, 256), mk(n, b, 256), mk(m, b, 64), mk(m, b, 64)
    go = torch.randn(n, b, 256, device=dev, generator=g)
    outs, grads = [], []
    for cast in (lambda t: t, lambda t: t.float()):
        c, mp = cast(curr).clone().requires_grad_(True), cast(mpos).clone().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=dtype):
            o = model(c, cast(mem), cast(cpos), mp, p)
        o.backward(go)
        outs.append(o.detach())
        grads.append((c.grad, mp.grad, torch.cat([q.grad.flatten() for q in model.parameters()])))
    assert outs[0].dtype == torch.float32 and torch.equal(outs[0], outs[1])
    assert grads[0][0].dtype == dtype and grads[0][1].dtype == dtype and grads[1][0].dtype == torch.float32
    assert rel_l2(grads[0][0], grads[1][0]) < 6e-3 and rel_l2(grads[0][1], grads[1][1]) < 6e-3      # the rounding of the gradient to 16 bits
    assert rel_l2(grads[0][2], grads[1][2]) < 1e-5


def test_fused_and_composed_paths_agree(dev):
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    model = build_memory_attention().to(dev).eval()
    inp = {k: v.to(dev) for k, v in detgen.attention_inputs(8, 2, 2, 8).items()}
    outs = []
    for fused in (True, False):
        model.use_fused_stack = fused
        with torch.no_grad():
            outs.append(model(inp["curr"], inp["memory"], inp["curr_pos"], inp["memory_pos"], 8))
    assert rel_l2(outs[0], outs[1]) < 5e-3


def test_grad_bucket_direct_accumulation_and_graph_replay(dev):
    """With a GradBucket attached the fused backward accumulates into the bucket itself (autograd sees None for the
    parameters); two backward passes must leave exactly the sum of the gradients the plain path returns, the bf16
    weight mirror must follow an in-place parameter update, and CUDA-graph replay must give the same numbers."""
    from sam2_video_training_b200 import ddp
    from sam2_video_training_b200.graphs import GraphedMemoryAttention
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    torch.manual_seed(3)
    ref_model = build_memory_attention(dropout=0.0).to(dev).train()
    model = build_memory_attention(dropout=0.0).to(dev).train()
    model.load_state_dict(ref_model.state_dict())
    bucket = ddp.attach_grad_bucket(model)
    graphed = GraphedMemoryAttention(model)
    g = torch.Generator(device="cuda").manual_seed(5)
    n, b, m, p = 64, 3, 2 * 64 + 8, 8

    def inputs():
        return (torch.randn(n, b, 256, device=dev, generator=g), torch.randn(m, b, 64, device=dev, generator=g),
                torch.randn(n, b, 256, device=dev, generator=g) * 0.7,
                (torch.randn(m, b, 64, device=dev, generator=g) * 0.7).requires_grad_(pos_grad),
                torch.randn(n, b, 256, device=dev, generator=g))

    for step in range(3):            # steps 1, 2 run after an in-place parameter update (mirror refresh)
        pos_grad = step < 2          # step 2: NO input requires grad -- the parameter gradients must still arrive
        batches = [inputs() for _ in range(2)]
        ref_model.zero_grad(set_to_none=True)
        bucket.zero()
        ref_pos_grads, pos_grads, outs_ref, outs = [], [], [], []
        for runner, sink_out, sink_pos, mdl in ((ref_model, outs_ref, ref_pos_grads, ref_model), (graphed, outs, pos_grads, model)):
            for curr, mem, cpos, mpos, go in batches:
                mpos.grad = None
                o = runner(curr, mem, cpos, mpos, p)
                o.backward(go)
                sink_out.append(o.detach().clone())
                if pos_grad:
                    sink_pos.append(mpos.grad.detach().clone())
                del o
        torch.cuda.synchronize()
        for a, c in zip(outs, outs_ref):
            assert rel_l2(a, c) < 1e-6, "graph replay / mirror changed the forward"
        for a, c in zip(pos_grads, ref_pos_grads):
            assert rel_l2(a, c) < 1e-5
        ref_grads = {name: pr.grad for name, pr in ref_model.named_parameters()}
        for (name, pr), pm in zip(ref_model.named_parameters(), model.parameters()):
            assert pm.grad.data_ptr() >= bucket.flat.data_ptr(), name          # still a view of the bucket
            # the self-attention q / k / v bias gradients of the direct path are column sums of the bf16 dq | dk | dv (extra
            # accumulator columns of the stacked weight-gradient GEMM); the autograd path sums the fp32 accumulators in the
            # attention epilogues -- the two differ by the bf16 rounding of the summands, everything else is the same arithmetic
            if "self_attn" in name and name.endswith("_proj.bias") and "out_proj" not in name:
                # (the k bias gradient is identically 0 in exact arithmetic -- rows of dS sum to 0 -- so it is pure rounding noise:
                # measure all three against the layer's largest q / k / v bias gradient)
                scale = max(float(ref_grads[name.replace(nm, alt)].norm()) for nm in ("q_proj", "k_proj", "v_proj") if nm in name
                            for alt in ("q_proj", "k_proj", "v_proj"))
                err = float((pm.grad - pr.grad).norm()) / scale
                assert err < 5e-3, (step, name, err)
                continue
            assert rel_l2(pm.grad, pr.grad) < 2e-5, (step, name, rel_l2(pm.grad, pr.grad))
        with torch.no_grad():        # same in-place update on both models
            for pr, pm in zip(ref_model.parameters(), model.parameters()):
                pr.add_(0.01 * torch.sign(pr.grad)); pm.add_(0.01 * torch.sign(pr.grad))


def _keep_mask(dev, seed, site, p_drop, n):
    """The keep mask the kernels derive from (seed, site) for element indices [0, n) (sam2b200_dropout_mask)."""
    from sam2_video_training_b200 import _lib
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    rc = _lib.load().sam2b200_dropout_mask(out.data_ptr(), 0, n, float(p_drop), seed.data_ptr(), int(site),
                                           torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "sam2b200_dropout_mask")
    return out.bool()


@pytest.mark.parametrize("b,n,m", [(2, 144, 296), (1, 1024, 1100)])
def test_attention_dropout_matches_explicit_mask(dev, b, n, m):
    """Attention-probability dropout (transformer.py:304-306) inside the forward and all backward kernels (separate
    kernels at N = 144, the CTA-pair kernel at N = 1024): same result as softmax -> explicit mask / (1 - p) -> @ V in
    fp32 torch with the mask the kernels generate, gradients included; keep rate ~ 1 - p."""
    from sam2_video_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(33)
    p_drop, site = 0.1, 11
    seed = torch.tensor([0x1234567890ABCDE], dtype=torch.int64, device=dev)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    keep = _keep_mask(dev, seed, site, p_drop, b * n * m).view(b, n, m)
    assert abs(float(keep.float().mean()) - (1 - p_drop)) < 5e-3
    drop = (p_drop, seed, site)
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0, drop=drop)
    dq, dk, dv = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, grad_dtype=torch.float32, drop=drop)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    a = torch.softmax(qf @ kf.transpose(1, 2) / 16.0, dim=-1) * keep / (1 - p_drop)
    ref = a @ vf
    ref.backward(do.float())
    assert rel_l2(o32, ref) < 5e-3
    for got, want, name in ((dq, qf.grad, "dq"), (dk, kf.grad, "dk"), (dv, vf.grad, "dv")):
        assert rel_l2(got, want) < 1e-2, (name, rel_l2(got, want))
    o_nodrop, _, _ = ops.attn_fwd(q, k, v, 1 / 16.0)
    assert rel_l2(o_nodrop, ref) > 0.05        # the mask really changes the result


@pytest.mark.parametrize("b", [2, 64])      # 64 objects: the cross-attention runs on the raw 64-d memory features
def test_training_mode_dropout_fused_stack_matches_oracle_with_same_masks(dev, b):
    """Train mode with the shipped dropout = 0.1 runs the fused stack; every dropout of the reference
    (memory_attention.py:64,81,97,99, transformer.py:304-306) is applied with counter-based masks.  The oracle, given
    the very masks the kernels generate, must reproduce output and gradients; two calls draw different masks."""
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    p_drop = 0.1
    model = build_memory_attention(dropout=p_drop).to(dev).train()
    assert model._fused_eligible()
    params = ao.init_params(seed=0)
    _load_params(model, params)
    grid, nf, nptr = 8, 2, 4
    n, m = grid * grid, nf * grid * grid + nptr
    g = torch.Generator().manual_seed(5)
    curr, curr_pos = torch.randn(n, b, 256, generator=g), torch.randn(n, b, 256, generator=g) * 0.7
    memory, memory_pos = torch.randn(m, b, 64, generator=g), torch.randn(m, b, 64, generator=g) * 0.7
    gout = torch.randn(n, b, 256, generator=g)
    seed_val = 987654321012345
    model._sam2b200_fixed_seed = seed_val
    cd = curr.to(dev).requires_grad_(True)
    out = model(cd, memory.to(dev), curr_pos.to(dev), memory_pos.to(dev), nptr)
    out.backward(gout.to(dev))
    seed = torch.tensor([seed_val], dtype=torch.int64, device=dev)
    masks = []
    for l in range(4):
        mk = lambda site, cnt, shape: _keep_mask(dev, seed, l * 8 + site, p_drop, cnt).view(shape).cpu()
        masks.append(dict(sa_prob=mk(0, b * n * n, (b, n, n)), ca_prob=mk(1, b * n * m, (b, n, m)),
                          drop1=mk(2, b * n * 256, (b, n, 256)), drop2=mk(3, b * n * 256, (b, n, 256)),
                          mlp=mk(4, b * n * 2048, (b, n, 2048)), drop3=mk(5, b * n * 256, (b, n, 256))))
    po = {k_: v_.clone().requires_grad_(True) for k_, v_ in params.items()}
    co = curr.clone().requires_grad_(True)
    ref = ao.memory_attention(po, co, memory, curr_pos, memory_pos, nptr, masks=masks, p_drop=p_drop)
    ref.backward(gout)
    assert rel_l2(out, ref) < ATTN_REL_TOL, rel_l2(out, ref)
    assert cosine(cd.grad, co.grad) > GRAD_COS_TOL
    got = torch.cat([p_.grad.flatten().cpu() for _, p_ in model.named_parameters()])
    want = torch.cat([po[k_].grad.flatten() for k_, _ in model.named_parameters()])
    assert cosine(got, want) > GRAD_COS_TOL, cosine(got, want)
    ref_nodrop = ao.memory_attention(params, curr, memory, curr_pos, memory_pos, nptr)
    assert rel_l2(out, ref_nodrop) > 0.05
    model._sam2b200_fixed_seed = None
    with torch.no_grad():
        o1 = model(curr.to(dev), memory.to(dev), curr_pos.to(dev), memory_pos.to(dev), nptr)
        o2 = model(curr.to(dev), memory.to(dev), curr_pos.to(dev), memory_pos.to(dev), nptr)
    assert rel_l2(o1, o2) > 0.02                # fresh masks per call
    model.eval()
    with torch.no_grad():
        o3 = model(curr.to(dev), memory.to(dev), curr_pos.to(dev), memory_pos.to(dev), nptr)
    assert rel_l2(o3, ref_nodrop) < ATTN_REL_TOL   # eval: dropout off


def test_dropout_under_cuda_graph_replay_draws_fresh_masks(dev):
    """The per-call seed is produced on the device inside the captured graph, so every replay has new masks, and the
    backward of a replay uses the masks of its own forward (gradient of a linear probe matches a finite difference
    of the SAME replay is not available -- instead: two replays differ, and eval-mode replay is deterministic)."""
    from sam2_video_training_b200 import ddp
    from sam2_video_training_b200.graphs import GraphedMemoryAttention
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    torch.manual_seed(1)
    model = build_memory_attention(dropout=0.1).to(dev).train()
    ddp.attach_grad_bucket(model)
    gm = GraphedMemoryAttention(model)
    n, b, m, p = 64, 2, 136, 8
    g = torch.Generator(device="cuda").manual_seed(2)
    curr, mem = torch.randn(n, b, 256, device=dev, generator=g), torch.randn(m, b, 64, device=dev, generator=g)
    cpos = torch.randn(n, b, 256, device=dev, generator=g)
    mpos = torch.randn(m, b, 64, device=dev, generator=g).requires_grad_(True)
    outs, grads = [], []
    for _ in range(3):
        mpos.grad = None
        o = gm(curr, mem, cpos, mpos, p)
        o.backward(torch.ones_like(o))
        outs.append(o.detach().clone()); grads.append(mpos.grad.detach().clone())
        del o
    assert len(gm._graphs) == 1
    assert rel_l2(outs[0], outs[1]) > 0.02 and rel_l2(outs[1], outs[2]) > 0.02
    assert rel_l2(grads[0], grads[1]) > 0.02
    assert all(torch.isfinite(t).all() for t in outs + grads)


# ------------------------------------------------------------------ glue kernels (csrc/glue.cu)
def test_ln_fwd_bwd_kernels(dev):
    from sam2_video_training_b200 import fused_stack as fs
    torch.manual_seed(3)
    rows, bq, nq = 6 * 50, 6, 50
    x = torch.randn(rows, 256, device=dev) * 2 + 0.3
    res = (torch.randn(rows, 256, device=dev)).to(torch.bfloat16)
    gamma = torch.randn(256, device=dev) * 0.2 + 1
    beta = torch.randn(256, device=dev) * 0.1
    y, xn, mean, rstd = fs.ln_fwd(x, res, gamma, beta)
    xr = (x + res.float()).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (256,), gr, br, 1e-5)
    assert torch.allclose(xn, xr.detach(), atol=1e-6)
    assert rel_l2(y, yr) < 4e-3 and y.dtype == torch.bfloat16
    dy = torch.randn(rows, 256, device=dev).to(torch.bfloat16)
    gin = torch.randn(rows, 256, device=dev)
    yr.backward(dy.float())
    dg, db = torch.zeros(256, device=dev), torch.zeros(256, device=dev)
    gout = fs.ln_bwd(dy, xn, mean, rstd, gamma, gin, dg, db)
    assert rel_l2(gout - gin, xr.grad) < 1e-5
    assert rel_l2(dg, gr.grad) < 1e-5 and rel_l2(db, br.grad) < 1e-5
    # fp32 seq-first output / transposed fp32 dy (final norm)
    y32, _, mean2, rstd2 = fs.ln_fwd(xn, None, gamma, beta, want_f32_seq_first=(bq, nq))
    assert y32.shape == (nq, bq, 256)
    assert torch.allclose(y32, yr.detach().view(bq, nq, 256).transpose(0, 1), atol=2e-5)
    dy32 = torch.randn(nq, bq, 256, device=dev)
    dg.zero_(); db.zero_()
    g2 = fs.ln_bwd(dy32, xn, mean2, rstd2, gamma, None, dg, db, seq_first=(bq, nq))
    xr2 = xn.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr2, (256,), gamma, beta, 1e-5).backward(dy32.transpose(0, 1).reshape(rows, 256))
    assert rel_l2(g2, xr2.grad) < 1e-5
    # fused bf16 copy + bias gradient of the next projection (column sums of the bf16-rounded output gradient)
    dg.zero_(); db.zero_()
    dbias = torch.full((256,), 0.25, device=dev)
    gout2, g16 = fs.ln_bwd(dy, xn, mean, rstd, gamma, gin, dg, db, dbias=dbias)
    assert torch.equal(gout2, gout) and g16.dtype == torch.bfloat16 and torch.equal(g16, gout.to(torch.bfloat16))
    assert rel_l2(dbias, g16.float().sum(0) + 0.25) < 1e-5
    assert rel_l2(dg, gr.grad) < 1e-5 and rel_l2(db, br.grad) < 1e-5


@pytest.mark.parametrize("k,n_out,b,length,n_rope", [(256, 3, 3, 64, 64), (256, 1, 2, 144, 144), (64, 1, 2, 136, 128), (64, 1, 3, 100, 0)])
def test_proj_rope_kernel(dev, k, n_out, b, length, n_rope):
    """sam2b200_proj_rope: x @ W^T + bias with the axial rotation fused into the GEMM epilogue, against fp32 torch on
    the same bf16 operands + the oracle's rotation (position_encoding.py:212-239): stacked q|k|v (v not rotated),
    memory keys with un-rotated object-pointer rows, ragged row counts."""
    from sam2_video_training_b200 import fused_stack as fs
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(k + n_out + length)
    grid = 12 if n_rope == 144 else 8            # n_rope is a multiple of the table period (keys tile the table)
    period = grid * grid
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    x = torch.randn(b * length, k, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(256 * n_out, k, device=dev, generator=g) / k ** 0.5).to(torch.bfloat16)
    bias = (torch.randn(256 * n_out, device=dev, generator=g) * 0.1).to(torch.bfloat16)
    rope_outs = 0 if n_rope == 0 else (2 if n_out == 3 else 1)
    outs = fs.proj_rope(x, w, bias, n_out, table if rope_outs else None, rope_outs, length, n_rope)
    ref = (x.float() @ w.float().t() + bias.float()).view(b, length, 256 * n_out).cpu()
    cos, sin = ao.axial_rope_table(period)
    for i, o in enumerate(outs):
        want = ref[:, :, 256 * i:256 * (i + 1)]
        if i < rope_outs:
            want = torch.cat([ao.apply_axial_rope(want[:, :n_rope], cos, sin), want[:, n_rope:]], dim=1)
        assert o.shape == (b * length, 256) and o.dtype == torch.bfloat16
        assert rel_l2(o.view(b, length, 256), want) < 4e-3, (i, rel_l2(o.view(b, length, 256), want))


@pytest.mark.parametrize("mode,b,length,with_res,drop", [("qkv", 3, 64, True, 0.0), ("q", 2, 144, True, 0.0), ("q", 1, 100, False, 0.0),
                                                         ("mlp", 2, 200, True, 0.0), ("mlp", 3, 70, True, 0.1), ("qkv", 2, 576, True, 0.1)])
def test_ln_proj_kernel(dev, mode, b, length, with_res, drop):
    """sam2b200_ln_proj: x + dropout(res) -> LayerNorm -> projection (+ bias) -> RoPE | ReLU (+ dropout) in one kernel, against
    the op sequence it replaces restated in fp32 torch on the same bf16-rounded operands (with the kernel's own dropout
    masks): x_new / mean / rstd to fp32 round-off, y and the outputs to bf16 rounding; ragged last row tile."""
    from sam2_video_training_b200 import fused_stack as fs
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(11)
    r = b * length
    grid = int(round(math.sqrt(length)))
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev) if grid * grid == length else None
    n_rope = length if table is not None else 0
    if table is None:      # a table for a smaller square grid: positions beyond it stay un-rotated
        grid = int(math.sqrt(length))
        table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
        n_rope = grid * grid
    x = torch.randn(r, 256, device=dev, generator=g) * 1.3 + 0.2
    res = torch.randn(r, 256, device=dev, generator=g).to(torch.bfloat16) if with_res else None
    gamma = 1 + 0.1 * torch.randn(256, device=dev, generator=g)
    beta = 0.1 * torch.randn(256, device=dev, generator=g)
    nout, n_out, width, rope_outs, relu = {"qkv": (768, 3, 256, 2, False), "q": (256, 1, 256, 1, False), "mlp": (2048, 1, 2048, 0, True)}[mode]
    w = (torch.randn(nout, 256, device=dev, generator=g) / 16).to(torch.bfloat16)
    bias = (torch.randn(nout, device=dev, generator=g) * 0.2).to(torch.bfloat16)
    seed = torch.tensor([987654321], dtype=torch.int64, device=dev)
    d_res = (drop, seed, 7) if drop > 0 and with_res else None
    d_out = (drop, seed, 12) if drop > 0 and relu else None
    outs, y, x_new, mean, rstd = fs.ln_proj(x, res, gamma, beta, w, bias, n_out, out_width=width, table=table if rope_outs else None,
                                            rope_outs=rope_outs, rows_per_item=length, n_rope_rows=n_rope, relu=relu,
                                            drop_res=d_res, drop_out=d_out)
    torch.cuda.synchronize()
    # ---- reference
    xr = x.clone()
    if with_res:
        rr = res.float()
        if d_res is not None:
            rr = rr * _keep_mask(dev, seed, 7, drop, r * 256).view(r, 256) / (1 - drop)
        xr = xr + rr
        assert rel_l2(x_new, xr) < 1e-6
    else:
        assert x_new.data_ptr() == x.data_ptr()
    mu = xr.mean(-1)
    var = ((xr - mu[:, None]) ** 2).mean(-1)
    assert rel_l2(mean, mu) < 1e-5 and rel_l2(rstd, torch.rsqrt(var + 1e-5)) < 1e-5
    yr = ((xr - mu[:, None]) * torch.rsqrt(var + 1e-5)[:, None] * gamma + beta)
    assert rel_l2(y, yr) < 4e-3
    o = y.float() @ w.float().t() + bias.float()           # the kernel's own bf16 y: the error budget below is the GEMM + epilogue only
    if rope_outs:
        o3 = o.view(b, length, nout)
        parts = []
        for i in range(n_out):
            blk = o3[:, :, i * 256:(i + 1) * 256]
            if i < rope_outs:
                cos, sin = ao.axial_rope_table(grid * grid)
                rot = ao.apply_axial_rope(blk[:, :n_rope].cpu(), cos, sin).to(dev)
                blk = torch.cat([rot, blk[:, n_rope:]], dim=1)
            parts.append(blk.reshape(r, 256))
    else:
        o = torch.relu(o)
        if d_out is not None:
            o = o * _keep_mask(dev, seed, 12, drop, r * nout).view(r, nout) / (1 - drop)
        parts = [o]
    for got, want in zip(outs, parts):
        assert got.shape == want.shape and got.dtype == torch.bfloat16
        assert rel_l2(got, want) < 4e-3, (mode, rel_l2(got, want))


@pytest.mark.parametrize("r,mo,no", [(200, 256, 256), (4096, 256, 64), (1000, 768, 256), (3000, 256, 2048), (2500, 2048, 256), (33, 256, 64),
                                    (32256, 256, 256)])
def test_wgrad_kernel(dev, r, mo, no):
    """sam2b200_wgrad: c += a^T b (bf16 operands, fp32 accumulation, split over the rows, partial tiles added with fp32
    reductions) against an fp32 GEMM on the same bf16 operands -- ragged row counts, strided operands (column slices of wider
    buffers, as the stacked q|k|v gradient is), accumulation into a non-zero c."""
    from sam2_video_training_b200 import fused_stack as fs
    g = torch.Generator(device="cuda").manual_seed(r + mo + no)
    a_wide = torch.randn(r, mo + 64, device=dev, generator=g).to(torch.bfloat16)
    b_wide = torch.randn(r, no + 128, device=dev, generator=g).to(torch.bfloat16)
    a, b = a_wide[:, 64:], b_wide[:, :no]
    c0 = torch.randn(mo, no, device=dev, generator=g)
    c = c0.clone()
    fs.wgrad_(c, a, b)
    torch.cuda.synchronize()
    want = c0.double() + a.double().t() @ b.double()
    assert rel_l2(c, want) < 2e-5, rel_l2(c, want)
    if no == 64 or (no == 256 and mo <= 768):
        # bias gradient from the same MMAs: 16 extra accumulator columns against a constant operand of ones
        c = c0.clone()
        db0 = torch.randn(mo, device=dev, generator=g)
        db = db0.clone()
        fs.wgrad_(c, a, b, db)
        torch.cuda.synchronize()
        assert rel_l2(c, want) < 2e-5, rel_l2(c, want)
        want_b = db0.double() + a.double().sum(0)
        assert rel_l2(db, want_b) < 2e-5, rel_l2(db, want_b)


@pytest.mark.parametrize("r", [200, 4096, 56 * 4060])
def test_wgrad_two_products_in_one_pass(dev, r):
    """sam2b200_wgrad2: c += a^T b and c2 += a^T b2 (+ column sums of a) from ONE pass over the long operand a [R, 256] -- the key
    projection's weight gradient together with the per-segment sums of the key gradient (one-hot b2) -- against fp64 products."""
    from sam2_video_training_b200 import fused_stack as fs
    g = torch.Generator(device="cuda").manual_seed(r)
    a = torch.randn(r, 256, device=dev, generator=g).to(torch.bfloat16)
    b = torch.randn(r, 64, device=dev, generator=g).to(torch.bfloat16)
    seg = torch.randint(0, 23, (r,), device=dev, generator=g)
    b2 = torch.nn.functional.one_hot(seg, 64).to(torch.bfloat16)
    c0, c20, db0 = (torch.randn(256, 64, device=dev, generator=g), torch.randn(512, 64, device=dev, generator=g)[128:384],
                    torch.randn(256, device=dev, generator=g))
    c, c2buf, db = c0.clone(), torch.zeros(512, 64, device=dev), db0.clone()
    c2 = c2buf[128:384]
    c2.copy_(c20)
    fs.wgrad2_(c, c2, a, b, b2, db)
    torch.cuda.synchronize()
    assert rel_l2(c, c0.double() + a.double().t() @ b.double()) < 2e-5
    assert rel_l2(c2, c20.double() + a.double().t() @ b2.double()) < 2e-5
    assert rel_l2(db, db0.double() + a.double().sum(0)) < 2e-5
    assert float(c2buf[:128].abs().max()) == 0.0 and float(c2buf[384:].abs().max()) == 0.0       # canary rows around the strided output


@pytest.mark.parametrize("r,k,no,nn", [(200, 256, 256, True), (32256, 256, 256, True), (1000, 768, 256, True), (3000, 2048, 256, True),
                                       (2500, 2048, 256, False), (40000, 256, 64, True), (33, 64, 256, False), (700, 256, 64, False),
                                       (19000, 64, 256, False)])
def test_gemm_kernel(dev, r, k, no, nn):
    """sam2b200_gemm (resident-CTA SS-mode tcgen05 GEMM, both weight layouts, fp32 bias) against an fp64 GEMM on the same bf16
    operands: ragged row counts (TMA clipping), more tiles than SMs (accumulator / ring phases across tiles), strided operands."""
    from sam2_video_training_b200 import fused_stack as fs
    g = torch.Generator(device="cuda").manual_seed(r + k + no)
    a_wide = torch.randn(r, k + 64, device=dev, generator=g).to(torch.bfloat16)
    a = a_wide[:, 64:]
    w = (torch.randn((k, no) if nn else (no, k), device=dev, generator=g) / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn(no, device=dev, generator=g) if not nn else None
    out = fs.gemm(a, w, nn=nn, bias=bias)
    torch.cuda.synchronize()
    want = a.double() @ (w.double() if nn else w.double().t())
    if bias is not None:
        want = want + bias.double()
    assert out.shape == (r, no) and out.dtype == torch.bfloat16
    assert rel_l2(out, want) < 3e-3, rel_l2(out, want)
    assert (out.double() - want).abs().max() < 0.05 * want.abs().max()
    if nn and no == 64:      # row dot products with an fp32 companion from the epilogue (Delta of the raw-memory attention backward)
        rows32 = torch.randn(r, 64, device=dev, generator=g)
        out2, dot = fs.gemm(a, w, nn=True, dot_rows=rows32)
        torch.cuda.synchronize()
        assert torch.equal(out2, out)
        want_dot = (out.double() * rows32.double()).sum(-1)
        assert rel_l2(dot, want_dot) < 1e-5, rel_l2(dot, want_dot)


@pytest.mark.parametrize("r,k,no,layout", [(200, 256, 256, 1), (33, 64, 256, 0), (700, 256, 64, 1), (2500, 2048, 256, 0), (19000, 256, 256, 0)])
def test_gemm_kernel_writes_stay_inside_the_output(dev, r, k, no, layout):
    """Bounds check without compute-sanitizer (closed on this pool): the C ABI writes into a caller-owned buffer with a wider row
    stride and guard rows; TMA clipping of the ragged last row tile must leave every guard element (columns >= No, rows >= R) untouched."""
    from sam2_video_training_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(r + k)
    a = torch.randn(r, k, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn((k, no) if layout else (no, k), device=dev, generator=g) / k ** 0.5).to(torch.bfloat16)
    ld, guard = no + 64, 160
    buf = torch.full((r + guard, ld), 7.0, dtype=torch.bfloat16, device=dev)
    rc = lib.sam2b200_gemm(buf.data_ptr(), ld, a.data_ptr(), k, w.data_ptr(), w.stride(0), layout, r, k, no, None, None, 1, 0, 1, None, None,
                           torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "sam2b200_gemm")
    torch.cuda.synchronize()
    want = a.double() @ (w.double() if layout else w.double().t())
    assert rel_l2(buf[:r, :no], want) < 3e-3
    assert bool((buf[:r, no:] == 7.0).all()) and bool((buf[r:] == 7.0).all())


@pytest.mark.parametrize("b,length,n_rope,grid,k", [(2, 136, 128, 8, 64), (3, 100, 0, 8, 64), (5, 4060 // 5, 576, 24, 64), (2, 144, 144, 12, 768),
                                                   (3, 4096 + 8, 4096, 64, 256)])
def test_gemm_kernel_memory_key_projection_with_rope(dev, b, length, n_rope, grid, k):
    """sam2b200_gemm as the memory-key projection (transformer.py:278, K = 64) with bias and the axial rotation on the fp32
    accumulator; object-pointer rows (position >= n_rope) stay un-rotated (transformer.py:296-302); keys tile the table."""
    from sam2_video_training_b200 import fused_stack as fs
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(length)
    period = grid * grid
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    # k = 64: the memory keys; k = 768: a rotated output with STREAMED weights (not used by the stack, allowed by the ABI: the layout
    # must leave room for the table cache); k = 256 on a 64 x 64 grid: the largest cached table (1024 px)
    x = torch.randn(b * length, k, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(256, k, device=dev, generator=g) / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn(256, device=dev, generator=g) * 0.1
    out = fs.gemm(x, w, bias=bias, table=table if n_rope else None, rows_per_item=length, n_rope_rows=n_rope)
    ref = (x.float() @ w.float().t() + bias).view(b, length, 256).cpu()
    cos, sin = ao.axial_rope_table(period)
    if n_rope:
        rot = ao.apply_axial_rope(ref[:, :n_rope], cos, sin)
        ref = torch.cat([rot, ref[:, n_rope:]], dim=1)
    assert rel_l2(out.view(b, length, 256), ref) < 4e-3, rel_l2(out.view(b, length, 256), ref)


@pytest.mark.parametrize("mode,b,length,drop", [("qkv", 3, 64, 0.0), ("q", 2, 144, 0.0), ("q", 1, 100, 0.0), ("mlp", 2, 200, 0.0),
                                                ("mlp", 3, 70, 0.1), ("qkv", 56, 576, 0.0), ("mlp", 56, 576, 0.1)])
def test_block_head_two_kernel_path_matches_fused_kernel_and_reference(dev, mode, b, length, drop):
    """ln_fwd + sam2b200_gemm_ex (column blocks, three outputs, RoPE / ReLU + dropout epilogues) against fp32 torch on the same bf16
    operands with the kernel's own dropout mask, and against the one-kernel sam2b200_ln_proj (same y / x_new / statistics bit for bit)."""
    from sam2_video_training_b200 import fused_stack as fs
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(17)
    r = b * length
    grid = {64: 8, 144: 12, 100: 10, 576: 24}.get(length, 0)
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev) if grid else None
    n_out, width, ropes, relu = {"qkv": (3, 256, 2, False), "q": (1, 256, 1, False), "mlp": (1, 2048, 0, True)}[mode]
    nout = n_out * width
    x = torch.randn(r, 256, device=dev, generator=g)
    res = torch.randn(r, 256, device=dev, generator=g).to(torch.bfloat16)
    gamma, beta = torch.rand(256, device=dev, generator=g) + 0.5, torch.randn(256, device=dev, generator=g) * 0.1
    w = (torch.randn(nout, 256, device=dev, generator=g) / 16).to(torch.bfloat16)
    bias = torch.randn(nout, device=dev, generator=g) * 0.1
    seed = torch.full((1,), 4321, dtype=torch.int64, device=dev)
    d_out = (drop, seed, 12) if drop > 0 else None
    kw = dict(out_width=width, table=table if ropes else None, rope_outs=ropes, rows_per_item=length, n_rope_rows=length, relu=relu, drop_out=d_out)
    outs, y, x_new, mean, rstd = fs.ln_then_proj(x, res, gamma, beta, w, bias, n_out, **kw)
    outs_f, y_f, x_new_f, mean_f, rstd_f = fs.ln_proj(x, res, gamma, beta, w, bias.to(torch.bfloat16), n_out, **kw)
    torch.cuda.synchronize()
    assert torch.equal(y, y_f) and torch.equal(x_new, x_new_f) and rel_l2(mean, mean_f) < 1e-6 and rel_l2(rstd, rstd_f) < 1e-6
    o = y.float() @ w.float().t() + bias
    if ropes:
        cos, sin = ao.axial_rope_table(grid * grid)
        parts = []
        for i in range(n_out):
            blk = o[:, 256 * i:256 * (i + 1)].view(b, length, 256).cpu()
            parts.append((ao.apply_axial_rope(blk, cos, sin) if i < ropes else blk).reshape(r, 256))
    else:
        if relu:
            o = torch.relu(o)
            if d_out is not None:
                o = o * _keep_mask(dev, seed, 12, drop, r * nout).view(r, nout) / (1 - drop)
        parts = [o[:, width * i:width * (i + 1)].cpu() for i in range(n_out)]
    for got, got_f, want in zip(outs, outs_f, parts):
        assert got.shape == want.shape and got.dtype == torch.bfloat16
        assert rel_l2(got, want) < 4e-3, (mode, rel_l2(got, want))
        assert rel_l2(got, got_f) < 4e-3, (mode, rel_l2(got, got_f))     # same operands; bias bf16 vs fp32, accumulation order


@pytest.mark.parametrize("rows", [128, 700, 4096])
def test_mlp_dh_kernel(dev, rows):
    """sam2b200_mlp_dh: dh = (dm @ W2) * (h > 0) * scale (tcgen05 GEMM, ReLU / hidden-dropout backward in the epilogue)
    against fp32 torch on the same bf16 operands; ragged row count, zeros of h exactly where the reference masks."""
    from sam2_video_training_b200 import fused_stack as fs
    g = torch.Generator(device="cuda").manual_seed(rows)
    dm = torch.randn(rows, 256, device=dev, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(256, 2048, device=dev, generator=g) / 16).to(torch.bfloat16)
    h = torch.relu(torch.randn(rows, 2048, device=dev, generator=g)).to(torch.bfloat16)
    for scale in (1.0, 1.0 / 0.9):
        dh = fs.mlp_dh(dm, w2, h, scale)
        ref = (dm.float() @ w2.float()) * (h.float() > 0) * scale
        assert dh.dtype == torch.bfloat16 and dh.shape == ref.shape
        assert rel_l2(dh, ref) < 4e-3, rel_l2(dh, ref)
        assert torch.equal(dh == 0, (ref.to(torch.bfloat16) == 0) | (h == 0))
        # the bias gradient of linear1 fused into the epilogue: column sums of the dh it stores, accumulated (+=)
        db = torch.full((2048,), 0.5, device=dev)
        dh2 = fs.mlp_dh(dm, w2, h, scale, dbias=db)
        assert torch.equal(dh2, dh)
        assert rel_l2(db - 0.5, dh.float().sum(0)) < 1e-4, rel_l2(db - 0.5, dh.float().sum(0))


@pytest.mark.parametrize("c", [256, 768, 2048])
def test_colsum_kernels(dev, c):
    from sam2_video_training_b200 import fused_stack as fs
    torch.manual_seed(4)
    rows = 777
    g32 = torch.randn(rows, c, device=dev)
    acc = torch.ones(c, device=dev)
    out16 = fs.cast_colsum(g32, acc)
    assert torch.equal(out16, g32.to(torch.bfloat16))
    assert torch.allclose(acc - 1, out16.float().sum(0), atol=2e-3, rtol=1e-4)
    x16 = torch.randn(rows, c, device=dev).to(torch.bfloat16)
    acc.zero_()
    fs.colsum_bf16(x16, acc)
    assert torch.allclose(acc, x16.float().sum(0), atol=2e-3, rtol=1e-4)
    h = torch.randn(rows, c, device=dev).to(torch.bfloat16)
    d = x16.clone()
    acc.zero_()
    fs.relu_bwd_colsum_(d, h, acc)
    ref = x16 * (h > 0)
    assert torch.equal(d, ref)
    assert torch.allclose(acc, ref.float().sum(0), atol=2e-3, rtol=1e-4)


def test_state_dict_roundtrip_and_eval_determinism(dev):
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    a = build_memory_attention().to(dev).eval()
    b = build_memory_attention().to(dev).eval()
    b.load_state_dict(a.state_dict(), strict=True)
    inp = {k: v.to(dev) for k, v in detgen.attention_inputs(8, 2, 2, 8).items()}
    with torch.no_grad():
        oa = a(inp["curr"], inp["memory"], inp["curr_pos"], inp["memory_pos"], 8)
        ob = b(inp["curr"], inp["memory"], inp["curr_pos"], inp["memory_pos"], 8)
    assert torch.equal(oa, ob)


# ------------------------------------------------------------------ losses
def _outs(logits, iou):
    t, c = logits.shape[:2]
    return [{"multistep_pred_multimasks_high_res": [logits[f]], "multistep_pred_ious": [iou[f]],
             "multistep_object_score_logits": [torch.zeros(c, 1, device=logits.device)]} for f in range(t)]


@pytest.mark.parametrize("tag", ["t2_c3_s16", "t3_c5_s40"])
def test_multistep_loss_vs_reference_golden(dev, golden_dir, tag):
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    g = np.load(os.path.join(golden_dir, f"loss_{tag}.npz"))
    t, c, s = int(g["t"]), int(g["c"]), int(g["s"])
    logits, targets, iou = detgen.loss_inputs(t, c, s)
    for mode, l1 in (("l1", True), ("mse", False)):
        crit = MultiStepMultiMasksAndIous(dict(W_FOCAL), supervise_all_iou=True, iou_use_l1_loss=l1)
        x = logits.to(dev).requires_grad_(True)
        ip = iou.to(dev).requires_grad_(True)
        out = crit(_outs(x, ip), targets.to(dev))
        out["total_loss"].backward()
        for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
            ref = float(g[f"{mode}:{k}"])
            assert abs(float(out[k]) - ref) <= LOSS_REL_TOL * abs(ref), (mode, k, float(out[k]), ref)
            assert out[k].dim() == 0
        assert float(out["loss_class"]) == 0.0
        assert rel_l2(x.grad, torch.from_numpy(g[f"{mode}:dlogits"])) < 1e-4
        assert rel_l2(ip.grad, torch.from_numpy(g[f"{mode}:diou"])) < 1e-5
    crit = MultiStepMultiMasksAndIous({"loss_mask": 1, "loss_dice": 10, "loss_iou": 10}, iou_use_l1_loss=True,
                                      logit_temperature=2.5, focal_alpha=0.6)
    x = logits.to(dev).requires_grad_(True)
    out = crit(_outs(x, iou.to(dev)), targets.to(dev))
    out["total_loss"].backward()
    assert abs(float(out["total_loss"]) - float(g["temp:total_loss"])) <= LOSS_REL_TOL * abs(float(g["temp:total_loss"]))
    assert rel_l2(x.grad, torch.from_numpy(g["temp:dlogits"])) < 1e-4


@pytest.mark.parametrize("tag", ["t2_c3_s16", "t3_c5_s40"])
def test_bce_loss_vs_reference_golden(dev, golden_dir, tag):
    from sam2_video_training_b200.losses import BCECategoryLoss
    g = np.load(os.path.join(golden_dir, f"loss_{tag}.npz"))
    t, c, s = int(g["t"]), int(g["c"]), int(g["s"])
    logits, targets, _ = detgen.loss_inputs(t, c, s)
    x = logits.to(dev).requires_grad_(True)
    out = BCECategoryLoss()([{"pred_masks_high_res": x[f]} for f in range(t)], targets.to(dev))
    out["total_loss"].backward()
    assert abs(float(out["total_loss"]) - float(g["bce:total_loss"])) <= LOSS_REL_TOL * abs(float(g["bce:total_loss"]))
    assert float(out["loss_bce"]) == float(out["total_loss"])
    assert rel_l2(x.grad, torch.from_numpy(g["bce:dlogits"])) < 1e-4
    tg = targets.clone()
    tg[:, :, 0, 0] = True
    pw = [1.5, 0.5, 2.0][:c] + [1.0] * max(0, c - 3)
    x = logits.to(dev).requires_grad_(True)
    out = BCECategoryLoss(pos_weight=pw, logit_temperature=1.7)([{"pred_masks": x[f]} for f in range(t)], tg.to(dev))
    out["total_loss"].backward()
    assert abs(float(out["total_loss"]) - float(g["bce_pw:total_loss"])) <= LOSS_REL_TOL * abs(float(g["bce_pw:total_loss"]))
    assert rel_l2(x.grad, torch.from_numpy(g["bce_pw:dlogits"])) < 1e-4


@pytest.mark.parametrize("t,c,s", [(10, 1, 384), (2, 7, 384), (1, 13, 512), (1, 3, 250)])
def test_multistep_loss_vs_oracle_real_shapes(dev, t, c, s):
    """BASELINE.json shapes (cfg1: 10 x 1 x 384^2; cfg2/3 channel counts), random logits ~ N(0, 4^2),
    elliptical targets, one empty channel where C allows; s = 250 exercises the non-vector path."""
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    g = torch.Generator().manual_seed(5 + t + c + s)
    logits = torch.randn(t, c, 1, s, s, generator=g) * 4
    yy, xx = torch.meshgrid(torch.arange(s), torch.arange(s), indexing="ij")
    targets = torch.zeros(t, c, s, s, dtype=torch.bool)
    for f in range(t):
        for ch in range(c):
            if c > 2 and ch == c - 1:
                continue
            cx, cy = 0.3 * s + 7 * ch, 0.5 * s - 3 * f
            targets[f, ch] = ((xx - cx) / (0.2 * s)) ** 2 + ((yy - cy) / (0.12 * s)) ** 2 < 1
    iou = torch.rand(t, c, 1, generator=g)
    x64 = logits.double().requires_grad_(True)
    i64 = iou.double().requires_grad_(True)
    ref = lo.multistep_loss([x64[f] for f in range(t)], targets, [i64[f] for f in range(t)], dict(W_FOCAL),
                            iou_use_l1_loss=True)
    ref["total_loss"].backward()
    crit = MultiStepMultiMasksAndIous(dict(W_FOCAL), supervise_all_iou=True, iou_use_l1_loss=True)
    x = logits.to(dev).requires_grad_(True)
    ip = iou.to(dev).requires_grad_(True)
    out = crit(_outs(x, ip), targets.to(dev))
    out["total_loss"].backward()
    for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
        assert abs(float(out[k]) - float(ref[k])) <= LOSS_REL_TOL * abs(float(ref[k])), k
    assert rel_l2(x.grad, x64.grad) < 1e-4
    assert cosine(x.grad, x64.grad) > 0.99999
    assert rel_l2(ip.grad, i64.grad) < 1e-5


def test_multistep_loss_multimask_vs_reference_golden(dev, golden_dir):
    """M = 3 masks per channel: the reference's valid filter flattens (channel, mask) pairs into rows (losses.py:149-166),
    reproduced on the same fused kernels; against the unmodified reference (tests/golden/loss_multimask_*.npz)."""
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    g = np.load(os.path.join(golden_dir, "loss_multimask_t2_c3_m3_s16.npz"))
    t, c, m, s = (int(g[k]) for k in ("t", "c", "m", "s"))
    logits, targets, iou_pred = detgen.multimask_loss_inputs(t, c, m, s)
    for mode, kw in (("l1_all", dict(iou_use_l1_loss=True, supervise_all_iou=True)), ("mse", dict(iou_use_l1_loss=False))):
        crit = MultiStepMultiMasksAndIous(dict(W_FOCAL), **kw)
        xs = [logits[f].to(dev).requires_grad_(True) for f in range(t)]
        ips = [iou_pred[f].to(dev).requires_grad_(True) for f in range(t)]
        outs = [{"multistep_pred_multimasks_high_res": [xs[f]], "multistep_pred_ious": [ips[f]],
                 "multistep_object_score_logits": [None]} for f in range(t)]
        got = crit(outs, targets.to(dev))
        got["total_loss"].backward()
        for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
            want = float(g[f"{mode}:{k}"])
            assert abs(float(got[k]) - want) <= 1e-5 * max(abs(want), 1.0), (mode, k, float(got[k]), want)
        assert rel_l2(torch.stack([x.grad for x in xs]), torch.from_numpy(g[f"{mode}:dlogits"])) < 1e-5
        assert rel_l2(torch.stack([x.grad for x in ips]), torch.from_numpy(g[f"{mode}:diou"])) < 1e-5


def test_loss_deferred_validity_check(dev):
    """check_valid="deferred": no host synchronisation per call; the "No valid masks" contract (losses.py:153-161) is
    enforced by raise_if_invalid(), which reads one device scalar -- or takes the value the caller already read."""
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    crit = MultiStepMultiMasksAndIous(dict(W_FOCAL), check_valid="deferred")
    x = torch.randn(2, 3, 1, 32, 32, device=dev)
    tg = torch.rand(2, 3, 32, 32, device=dev) > 0.5
    mk = lambda: [{"multistep_pred_multimasks_high_res": [x[f]], "multistep_pred_ious": [torch.rand(3, 1, device=dev)],
                   "multistep_object_score_logits": [None]} for f in range(2)]
    crit(mk(), tg)
    assert int(crit.deferred_state().item()) == 3
    crit.raise_if_invalid()                      # fine, and resets
    assert crit.deferred_state() is None
    crit(mk(), tg)
    bad = tg.clone()
    bad[1] = False                               # frame 1 has no foreground at all
    crit(mk(), bad)                              # does not raise here ...
    crit(mk(), tg)
    with pytest.raises(ValueError, match="No valid masks"):
        crit.raise_if_invalid()                  # ... but here, whatever came after
    crit(mk(), bad)
    with pytest.raises(ValueError, match="No valid masks"):
        crit.raise_if_invalid(value=int(crit.deferred_state().item()))


def test_loss_no_valid_masks_raises(dev):
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    logits, targets, iou = detgen.loss_inputs(2, 3, 16)
    targets[1] = False
    crit = MultiStepMultiMasksAndIous(dict(W_FOCAL))
    with pytest.raises(ValueError, match="No valid masks"):
        crit(_outs(logits.to(dev), iou.to(dev)), targets.to(dev))


def test_loss_frames_are_used_in_place(dev):
    """Per-frame logits that are separate allocations (as the training wrapper produces) give the
    same result as views of one stacked tensor."""
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    logits, targets, iou = detgen.loss_inputs(3, 4, 32, clear=())
    crit = MultiStepMultiMasksAndIous(dict(W_FOCAL))
    a = crit(_outs(logits.to(dev), iou.to(dev)), targets.to(dev))
    sep = [logits[f].to(dev).clone() for f in range(3)]
    outs = [{"multistep_pred_multimasks_high_res": [sep[f]], "multistep_pred_ious": [iou[f].to(dev)],
             "multistep_object_score_logits": [torch.zeros(4, 1, device=dev)]} for f in range(3)]
    b = crit(outs, targets.to(dev))
    assert float(a["total_loss"]) == float(b["total_loss"])


@pytest.mark.parametrize("n_clips,t,c,s", [(8, 10, 7, 96), (3, 30, 2, 40), (5, 30, 1, 24), (2, 3, 4, 250)])
def test_loss_over_several_clips_in_one_launch(dev, n_clips, t, c, s):
    """forward_clips: the frames of all clips of a step through ONE launch (one target pointer per frame; 80 frames = the
    cfg2 step, 90 and 150 frames = more than the old 64-frame table, 150 = two sub-launches) equals the sum of the per-clip calls
    -- loss values to fp32 round-off, gradients w.r.t. every frame's logits and IoU predictions bit for bit."""
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    g = torch.Generator().manual_seed(n_clips * 100 + t)
    crit = MultiStepMultiMasksAndIous(dict(W_FOCAL), supervise_all_iou=True, iou_use_l1_loss=True)
    clips, leaves = [], []
    for ci in range(n_clips):
        logits = (torch.randn(t, c, 1, s, s, generator=g) * 3).to(dev)
        targets = (torch.rand(t, c, s, s, generator=g) > 0.6).to(dev)
        iou = torch.rand(t, c, 1, generator=g).to(dev)
        xs = [logits[f].clone().requires_grad_(True) for f in range(t)]
        ips = [iou[f].clone().requires_grad_(True) for f in range(t)]
        outs = [{"multistep_pred_multimasks_high_res": [xs[f]], "multistep_pred_ious": [ips[f]],
                 "multistep_object_score_logits": [None]} for f in range(t)]
        clips.append((outs, targets))
        leaves.append((xs, ips))
    ref = None
    for outs, targets in clips:
        o = crit(outs, targets)
        o["total_loss"].backward()
        ref = {k: float(v) for k, v in o.items()} if ref is None else {k: ref[k] + float(v) for k, v in o.items()}
    want = [[x.grad.clone() for x in xs] + [i.grad.clone() for i in ips] for xs, ips in leaves]
    for xs, ips in leaves:
        for v in xs + ips:
            v.grad = None
    got = crit.forward_clips(clips)
    got["total_loss"].backward()
    for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
        assert abs(float(got[k]) - ref[k]) <= 2e-6 * abs(ref[k]), (k, float(got[k]), ref[k])
    for (xs, ips), w in zip(leaves, want):
        for v, wv in zip(xs + ips, w):
            assert torch.equal(v.grad, wv)
    # the validity contract covers every clip
    bad_t = clips[-1][1].clone()
    bad_t[t - 1] = False
    with pytest.raises(ValueError, match="No valid masks"):
        crit.forward_clips(clips[:-1] + [(clips[-1][0], bad_t)])
    with pytest.raises(ValueError):
        crit.forward_clips(clips[:1] + [(clips[1][0], clips[1][1][:, :, : s // 2])])


@pytest.mark.parametrize("multimask", [True, False])
def test_functional_losses_match_reference_formulas(dev, multimask):
    """dice_loss / sigmoid_focal_loss / iou_loss keep the reference's signatures and values (losses.py:20-76); the
    expected values come from the oracle's restatement of those three functions, gradients included."""
    from oracle import losses_oracle as lo
    from sam2_video_training_b200 import losses as L
    g = torch.Generator().manual_seed(21)
    n, m, s = 3, 2, 48
    x = (torch.randn(n, m, s, s, generator=g) * 3).double().requires_grad_(True)
    t = (torch.rand(n, m, s, s, generator=g) > 0.7).double()
    iou = torch.rand(n, m, generator=g).double().requires_grad_(True)
    xg = x.detach().float().to(dev).requires_grad_(True)
    tg = t.float().to(dev)
    ig = iou.detach().float().to(dev).requires_grad_(True)
    if multimask:
        ref = (lo.dice_loss(x, t, 5.0, True), lo.sigmoid_focal_loss(x, t, 5.0, 0.25, 2.0, True), lo.iou_loss(x, t, iou, 5.0, True, False))
        got = (L.dice_loss(xg, tg, 5.0, True), L.sigmoid_focal_loss(xg, tg, 5.0, 0.25, 2.0, True), L.iou_loss(xg, tg, ig, 5.0, True, False))
    else:
        xf, tf = x.flatten(1), t.flatten(1)
        ref = (lo.dice_loss(xf, tf, 5.0), lo.sigmoid_focal_loss(xf, tf, 5.0, 0.25, 2.0), lo.iou_loss(x, t, iou, 5.0, False, True))
        got = (L.dice_loss(xg.flatten(1), tg.flatten(1), 5.0), L.sigmoid_focal_loss(xg.flatten(1), tg.flatten(1), 5.0, 0.25, 2.0),
               L.iou_loss(xg, tg, ig, 5.0, False, True))
    w = [torch.rand(r.shape, generator=g).double() + 0.5 for r in ref]
    sum((r * wi).sum() for r, wi in zip(ref, w)).backward()
    sum((o * wi.float().to(dev)).sum() for o, wi in zip(got, w)).backward()
    for r, o in zip(ref, got):
        assert o.shape == r.shape
        assert rel_l2(o.detach().cpu().double(), r.detach()) < 1e-5
    assert rel_l2(xg.grad.cpu().double(), x.grad) < 1e-5
    assert rel_l2(ig.grad.cpu().double(), iou.grad) < 1e-5


def test_attention_backward_pair_kernel_long_query_loop(dev):
    """N >= 1024 takes the CTA-pair key-side kernel (dV + dK on clusters of 2, P^T through distributed shared memory):
    same gradients as an fp32 torch restatement of the attention core, ragged key count, rotated and un-rotated keys."""
    from sam2_video_training_b200 import ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(17)
    grid, b, nptr = 32, 2, 20
    n, m = grid * grid, 2 * grid * grid + nptr             # N = 1024, M = 2068 (last key block ragged)
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0)
    db = [torch.zeros(256, device=dev) for _ in range(3)]
    dq, dk, dv = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, grad_dtype=torch.float32, dbias=tuple(db))
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = torch.softmax(qf @ kf.transpose(1, 2) / 16.0, dim=-1) @ vf
    ref.backward(do.float())
    assert rel_l2(o32, ref) < 5e-3
    for got, want, name in ((dq, qf.grad, "dq"), (dk, kf.grad, "dk"), (dv, vf.grad, "dv")):
        assert rel_l2(got, want) < 8e-3, (name, rel_l2(got, want))
    # without RoPE the column sums of dK vanish analytically (sum_j dS_ij = 0): compare on the scale of sum |dK|
    assert float((db[1] - dk.double().sum((0, 1)).float()).abs().max()) < 1e-4 * float(dk.abs().sum((0, 1)).max())
    assert rel_l2(db[2], vf.grad.sum((0, 1))) < 8e-3
    # fused conjugate rotation in the pair kernel's dK epilogue == rotating the plain gradient back
    dq2, dk2, dv2 = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=m - nptr, grad_dtype=torch.float32)
    want = ops.rope_apply(dk, table, m - nptr, inverse=True, out_dtype=torch.float32)
    assert rel_l2(dk2, want) < 1e-5 and torch.equal(dv2, dv)


@pytest.mark.parametrize("b,grid,nf,extra,drop", [(8, 24, 4, 196, 0.0), (5, 24, 7, 28, 0.1), (40, 8, 7, 12, 0.0)])
def test_attention_backward_persistent_key_side_kernels(dev, b, grid, nf, extra, drop):
    """bf16 gradients with more (key block, object) items than SMs take the PERSISTENT dV / dK kernels
    (attn_persist_kernels.cuh: resident CTAs, slot ring with operand prefetch and an epilogue slot); fp32 gradients take
    the one-CTA-per-item kernels.  Per item both run the same MMAs in the same order, so the bf16 results must be the
    bf16 rounding of the fp32 results bit for bit -- ragged last key block, RoPE, dropout, fused bias gradients."""
    from sam2_video_training_b200 import ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(23)
    n = grid * grid
    m = nf * n + extra
    assert ((m + 127) // 128) * b > torch.cuda.get_device_properties(dev).multi_processor_count
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    dr = None
    if drop > 0:
        dr = (drop, torch.tensor([12345], dtype=torch.int64, device=dev), 3)
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0, drop=dr)
    res = {}
    for gdt in (torch.float32, torch.bfloat16):
        db = [torch.zeros(256, device=dev) for _ in range(3)]
        res[gdt] = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=gdt,
                                dbias=tuple(db), drop=dr) + (db,)
    torch.cuda.synchronize()
    for i, name in ((1, "dk"), (2, "dv")):
        assert torch.equal(res[torch.float32][i].to(torch.bfloat16), res[torch.bfloat16][i]), name
    for j in (1, 2):
        assert rel_l2(res[torch.bfloat16][3][j], res[torch.float32][3][j]) < 2e-3
    if drop == 0.0:   # and against an fp32 torch restatement of the core
        # q and k ARE the rotated operands; `table` only makes the kernels rotate the gradients back
        qf, kf, vf = q.float().requires_grad_(True), k.float().requires_grad_(True), v.float().requires_grad_(True)
        ref = torch.softmax(qf @ kf.transpose(1, 2) / 16.0, dim=-1) @ vf
        ref.backward(do.float())
        want_dk = ops.rope_apply(kf.grad, table, nf * n, inverse=True, out_dtype=torch.float32)
        assert rel_l2(res[torch.bfloat16][1].float(), want_dk) < 8e-3
        assert rel_l2(res[torch.bfloat16][2].float(), vf.grad) < 8e-3


@pytest.mark.parametrize("b,grid,nf,nptr", [(3, 12, 2, 8), (2, 24, 7, 28), (1, 32, 2, 20), (40, 8, 7, 12), (40, 12, 4, 16)])   # last two: more (key block, object) items than SMs, 1 and 3 query tiles
def test_cross_attention_on_raw_memory_features(dev, b, grid, nf, nptr):
    """sam2b200_attn_fwd_v64 / _bwd_v64: softmax(q k^T) (mem Wv^T + bv) == (softmax(q k^T) mem) Wv^T + bv (rows sum to 1),
    so the value projection moves from the [B M, 64] memory to the [B N, 64] result.  Output, dq, dk and the value
    projection's gradients against an fp32 torch restatement of the ORIGINAL formulation and against the 256-d kernels."""
    from sam2_video_training_b200 import ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(41)
    n = grid * grid
    m = nf * n + nptr
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    wv = (torch.randn(256, 64, device=dev, generator=g) / 8).to(torch.bfloat16)
    bv = torch.randn(256, device=dev, generator=g) * 0.1
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    # reference: the reference's formulation in fp32
    qf, kf, mf, wf, bf = (t.float().requires_grad_(True) for t in (q, k, mem, wv, bv))
    ref = torch.softmax(qf @ kf.transpose(1, 2) / 16.0, dim=-1) @ (mf @ wf.t() + bf)
    ref.backward(do.float())
    # new path
    o64, o64_32, lse, _ = ops.attn_fwd_v64(q, k, mem, 1 / 16.0)
    out = o64_32 @ wv.float().t() + bv
    assert rel_l2(out, ref) < 5e-3
    do64 = (do.float() @ wv.float()).to(torch.bfloat16)
    delta = (do64.float() * o64_32).sum(-1)
    dq, dk = ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16.0, grad_dtype=torch.float32)
    assert rel_l2(dq, qf.grad) < 1e-2, rel_l2(dq, qf.grad)
    assert rel_l2(dk, kf.grad) < 1e-2, rel_l2(dk, kf.grad)
    d_wv = do.float().flatten(0, 1).t() @ o64_32.flatten(0, 1)         # [256, 64]
    assert rel_l2(d_wv, wf.grad) < 5e-3
    assert rel_l2(do.float().sum((0, 1)), bf.grad) < 1e-5
    # the 256-d kernels on the projected values give the same thing
    v = (mem.float() @ wv.float().t() + bv).to(torch.bfloat16)
    o, o32, lse_old = ops.attn_fwd(q, k, v, 1 / 16.0)
    assert rel_l2(out, o32) < 5e-3 and float((lse - lse_old).abs().max()) < 1e-3
    # fused conjugate rotation + bias gradients in the epilogues, bf16 gradients, each part on its own
    db = [torch.zeros(256, device=dev) for _ in range(2)]
    dq2, _ = ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=torch.bfloat16,
                              dbias=(db[0], None), parts=8)
    _, dk2 = ops.attn_bwd_v64(q, k, mem, do64, lse, delta, 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=torch.bfloat16,
                              dbias=(None, db[1]), parts=4)
    want_q = ops.rope_apply(dq, table, n, inverse=True, out_dtype=torch.float32)
    want_k = ops.rope_apply(dk, table, nf * n, inverse=True, out_dtype=torch.float32)
    assert rel_l2(dq2.float(), want_q) < 6e-3 and rel_l2(dk2.float(), want_k) < 6e-3
    assert rel_l2(db[0], want_q.sum((0, 1))) < 5e-3 and rel_l2(db[1], want_k.sum((0, 1))) < 5e-3


@pytest.mark.parametrize("b,grid,nf,nptr,drop", [(3, 12, 2, 8, 0.0), (2, 24, 7, 28, 0.0), (1, 32, 3, 20, 0.0), (40, 8, 7, 12, 0.0),
                                                  (2, 16, 3, 12, 0.1), (1, 8, 1, 0, 0.0)])
def test_raw_memory_backward_two_softmax_groups_bit_identical(dev, b, grid, nf, nptr, drop):
    """The experimental three_gemm_v64x2_kernel (two softmax groups on alternate tiles, 16-warp epilogue; variant 1 of
    sam2b200_debug_set_variant key 0, not the default) runs the same products in the same order as the single-group kernel: dq / dk are bit-identical in fp32 and bf16 (ragged last tiles, odd and even tile
    counts, a single tile, RoPE, dropout); the bias gradients agree to fp32 summation order."""
    from sam2_video_training_b200 import _lib, ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(97)
    n = grid * grid
    m = nf * n + nptr
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
    dr = (drop, torch.tensor([4242], dtype=torch.int64, device=dev), 5) if drop > 0 else None
    o64, o32, lse, rs = ops.attn_fwd_v64(q, k, mem, 1 / 16.0, drop=dr)
    delta = (do64.float() * o32).sum(-1)
    c = torch.randn(b, n, device=dev, generator=g) if drop > 0 else None
    if drop > 0:
        delta = delta + c * rs
    res = {}
    try:
        for variant in (0, 1):
            lib.sam2b200_debug_set_variant(0, variant)
            for gdt in (torch.float32, torch.bfloat16):
                db = [torch.zeros(256, device=dev) for _ in range(2)]
                dq, dk = ops.attn_bwd_v64(q, k, mem, do64, lse, delta.contiguous(), 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=gdt,
                                          dbias=tuple(db), dp_bias=c, drop=dr)
                torch.cuda.synchronize()
                res[(variant, gdt)] = (dq, dk, db)
    finally:
        lib.sam2b200_debug_set_variant(0, 0)
    for gdt in (torch.float32, torch.bfloat16):
        one, two = res[(0, gdt)], res[(1, gdt)]
        assert torch.equal(one[0], two[0]), ("dq", gdt)
        assert torch.equal(one[1], two[1]), ("dk", gdt)
        for i in range(2):
            assert rel_l2(two[2][i], one[2][i]) < 1e-4, ("dbias", i, gdt)


@pytest.mark.parametrize("b,grid,nf,nptr,drop", [(56, 24, 2, 8, 0.0), (40, 24, 1, 4, 0.0), (150, 12, 3, 8, 0.0), (60, 16, 3, 12, 0.1), (9, 32, 2, 20, 0.0),
                                                  (200, 12, 1, 0, 0.1)])
def test_raw_memory_backward_persistent_bit_identical(dev, b, grid, nf, nptr, drop):
    """three_gemm_v64_persistent_kernel (default when there are more (row block, object) items than SMs: resident CTAs, the next
    item's operands fetched under the current epilogue, output staging in the idle tile ring) against one CTA per item
    (sam2b200_debug_set_variant key 0 = 2; 3 = resident CTAs for dK and dQ at every length): the same instructions on the same data, so dq / dk must be bit-identical -- odd and
    even tile counts, 1..13 items per CTA, ragged last row blocks and tiles, frame boundaries inside a row block, dropout."""
    from sam2_video_training_b200 import _lib, ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(1234 + b)
    n = grid * grid
    m = nf * n + nptr
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
    dr = (drop, torch.tensor([99], dtype=torch.int64, device=dev), 2) if drop > 0 else None
    o64, o32, lse, rs = ops.attn_fwd_v64(q, k, mem, 1 / 16.0, drop=dr)
    delta = (do64.float() * o32).sum(-1)
    c = torch.randn(b, n, device=dev, generator=g) if drop > 0 else None
    if drop > 0:
        delta = delta + c * rs
    res = {}
    try:
        for variant in (2, 3):
            lib.sam2b200_debug_set_variant(0, variant)
            db = [torch.zeros(256, device=dev) for _ in range(2)]
            dq, dk = ops.attn_bwd_v64(q, k, mem, do64, lse, delta.contiguous(), 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=torch.bfloat16,
                                      dbias=tuple(db), dp_bias=c, drop=dr)
            dq2, dk2 = ops.attn_bwd_v64(q, k, mem, do64, lse, delta.contiguous(), 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=torch.bfloat16,
                                        dp_bias=c, drop=dr)
            torch.cuda.synchronize()
            res[variant] = (dq, dk, dq2, dk2, db)
    finally:
        lib.sam2b200_debug_set_variant(0, 0)
    for i in range(4):
        assert torch.equal(res[3][i], res[2][i]), i
    for i in range(2):
        assert rel_l2(res[3][4][i], res[2][4][i]) < 1e-4, ("dbias", i)
    assert torch.isfinite(res[3][0].float()).all() and torch.isfinite(res[3][1].float()).all()


@pytest.mark.parametrize("b,n,m,drop", [(2, 576, 4060, 0.0), (3, 64, 64, 0.0), (2, 200, 129, 0.0), (1, 1024, 3092, 0.0), (2, 144, 300, 0.1),
                                        (5, 576, 128, 0.0), (1, 130, 1000, 0.1)])
def test_raw_memory_forward_two_softmax_streams(dev, b, n, m, drop):
    """fwd_v64x2_kernel (experiment, sam2b200_debug_set_variant key 2 = 1; two online-softmax streams on alternate memory tiles,
    merged at the end) against the default single-stream kernel and an fp32 torch softmax: one tile, two tiles, odd and even
    tile counts, ragged query and memory tiles, dropout (same Philox stream, so the same mask), with and without the fused
    output projection.  The row sums (lse2, rowsum) are fp32 sums of the same terms in another order; the outputs differ by the
    bf16 rounding of the probabilities, which are taken relative to each stream's own running maximum (2^-9 per term)."""
    from sam2_video_training_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(11 + n + m)
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = (0.5 * torch.randn(b, m, 256, device=dev, generator=g)).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    w16 = (0.1 * torch.randn(256, 64, device=dev, generator=g)).to(torch.bfloat16)
    bias = torch.randn(256, device=dev, generator=g)
    rank1 = torch.randn(256, device=dev, generator=g)
    dr = (drop, torch.tensor([77], dtype=torch.int64, device=dev), 3) if drop > 0 else None
    res = {}
    try:
        for variant in (0, 1):
            lib.sam2b200_debug_set_variant(2, variant)
            a = ops.attn_fwd_v64(q, k, mem, 1 / 16.0, drop=dr)
            c = ops.attn_fwd_v64_proj(q, k, mem, w16, bias, rank1 if dr is not None else None, 1 / 16.0, drop=dr)
            torch.cuda.synchronize()
            res[variant] = (a, c)
    finally:
        lib.sam2b200_debug_set_variant(2, 0)
    (one, one_p), (two, two_p) = res[0], res[1]
    for x in (two, two_p[:4]):
        assert rel_l2(x[1], one[1]) < 3e-3, "out32"
        assert (x[2] - one[2]).abs().max().item() < 1e-4 * max(1.0, one[2].abs().max().item()), "lse2"
        assert rel_l2(x[0].float(), one[0].float()) < 4e-3, "out bf16"
        if dr is not None:
            assert rel_l2(x[3], one[3]) < 1e-4, "rowsum"
    assert rel_l2(two_p[4].float(), one_p[4].float()) < 4e-3, "proj"
    if dr is None:
        s = torch.einsum("bnd,bmd->bnm", q.float(), k.float()) / 16.0
        want = torch.softmax(s, -1) @ mem.float()
        assert rel_l2(two[1], want) < 6e-3
        assert (two[2] - torch.logsumexp(s, -1) * 1.4426950408889634).abs().max().item() < 2e-3
        assert rel_l2(two_p[4].float(), want @ w16.float().t() + bias) < 8e-3


@pytest.mark.parametrize("grid", [8, 24, 32])
def test_gradient_epilogue_axial_table_addressing_bit_identical(dev, grid):
    """The gradient epilogues read the (cos, sin) pairs of position p from rows p mod w (x part) and p - p mod w (y part)
    of the axial table instead of row p (w + 128 / w rows per CTA instead of 128): the values are the same numbers, so the
    rotated-back gradients must be bit-identical to full-row addressing (sam2b200_debug_set_variant key 1)."""
    from sam2_video_training_b200 import _lib, ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    n, b = grid * grid, 2
    m = 2 * n + 12
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q, do = (torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
    k, v = (torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0)
    o64, o64_32, lse64, _ = ops.attn_fwd_v64(q, k, mem, 1 / 16.0)
    delta = (do64.float() * o64_32).sum(-1)
    res = {}
    try:
        for variant in (1, 0):
            lib.sam2b200_debug_set_variant(1, variant)
            a = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=2 * n, grad_dtype=torch.float32)
            c = ops.attn_bwd_v64(q, k, mem, do64, lse64, delta, 1 / 16.0, table=table, n_rope_k=2 * n, grad_dtype=torch.bfloat16)
            torch.cuda.synchronize()
            res[variant] = (*a, *c)
    finally:
        lib.sam2b200_debug_set_variant(1, 0)
    for x, y in zip(res[0], res[1]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("b,grid,nf,nptr", [(2, 8, 2, 12), (3, 24, 2, 12), (56, 24, 1, 0), (2, 32, 3, 20), (40, 12, 2, 8), (1, 64, 1, 4)])
def test_gradient_epilogue_staged_rotation_table_bit_identical(dev, b, grid, nf, nptr):
    """The gradient epilogues rotate with a copy of the CTA's table rows staged in shared memory at kernel start (default)
    instead of loading them from global memory per 32-column chunk (sam2b200_debug_set_variant key 3 = 1): same numbers, so dq / dk
    must be bit-identical -- one-shot and persistent key-side kernels (b * key blocks > 148), CTAs whose rows straddle a frame
    boundary and the pointer tokens, grids whose table does not fit next to the raw-memory kernels (64: falls back)."""
    from sam2_video_training_b200 import _lib, ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    n = grid * grid
    m = nf * n + nptr
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q, do = (torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
    k, v = (torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    do64 = torch.randn(b, n, 64, device=dev, generator=g).to(torch.bfloat16)
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0)
    o64, o64_32, lse64, _ = ops.attn_fwd_v64(q, k, mem, 1 / 16.0)
    delta = (do64.float() * o64_32).sum(-1)
    res = {}
    try:
        for variant in (1, 0):
            lib.sam2b200_debug_set_variant(3, variant)
            a = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=torch.bfloat16)
            c = ops.attn_bwd_v64(q, k, mem, do64, lse64, delta, 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=torch.bfloat16)
            a32 = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=nf * n, grad_dtype=torch.float32)
            torch.cuda.synchronize()
            res[variant] = (*a, *c, *a32)
    finally:
        lib.sam2b200_debug_set_variant(3, 0)
    for i, (x, y) in enumerate(zip(res[0], res[1])):
        assert torch.equal(x, y), i


@pytest.mark.parametrize("b,grid,nf,nptr,drop", [(3, 12, 2, 8, 0.0), (2, 24, 7, 28, 0.0), (40, 8, 3, 12, 0.0), (2, 16, 2, 4, 0.1)])
def test_attention_forward_with_fused_output_projection(dev, b, grid, nf, nptr, drop):
    """The output projection (transformer.py:308-309) inside the attention kernels' epilogue: sam2b200_attn_fwd_proj (256-d
    values, Wo) and sam2b200_attn_fwd_v64_proj (raw memory features, folded Wo Wv; with dropout the rank-1 row-sum term).  The
    un-projected outputs / lse must be BIT-IDENTICAL to the un-fused kernels; the projection is compared with an fp32 GEMM on
    the kernel's own bf16 output (the tile the tensor core multiplies) -- ragged last query tile, several key tiles."""
    from sam2_video_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(55)
    n = grid * grid
    m = nf * n + nptr
    q = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    mem = torch.randn(b, m, 64, device=dev, generator=g).to(torch.bfloat16)
    wo = (torch.randn(256, 256, device=dev, generator=g) / 16).to(torch.bfloat16)
    weff = (torch.randn(256, 64, device=dev, generator=g) / 8).to(torch.bfloat16)
    bias = torch.randn(256, device=dev, generator=g) * 0.3
    rank1 = torch.randn(256, device=dev, generator=g) * 0.3
    dr = (drop, torch.tensor([777], dtype=torch.int64, device=dev), 6) if drop > 0 else None
    # 256-d values
    o0, o0_32, lse0 = ops.attn_fwd(q, k, v, 1 / 16.0, 1, drop=dr)
    o1, o1_32, lse1, proj = ops.attn_fwd_proj(q, k, v, wo, bias, 1 / 16.0, drop=dr)
    torch.cuda.synchronize()
    assert torch.equal(o0, o1) and torch.equal(lse0, lse1) and torch.equal(o0_32, o1_32)
    want = o1.float() @ wo.float().t() + bias
    assert rel_l2(proj, want) < 4e-3, rel_l2(proj, want)
    # raw memory features + folded projection
    a0 = ops.attn_fwd_v64(q, k, mem, 1 / 16.0, drop=dr)
    a1 = ops.attn_fwd_v64_proj(q, k, mem, weff, bias, rank1 if dr is not None else None, 1 / 16.0, drop=dr)
    torch.cuda.synchronize()
    assert torch.equal(a0[0], a1[0]) and torch.equal(a0[1], a1[1]) and torch.equal(a0[2], a1[2])
    want = a1[0].float() @ weff.float().t() + bias
    if dr is not None:
        assert torch.equal(a0[3], a1[3])
        want = want + a1[3][..., None] * rank1
    assert rel_l2(a1[4], want) < 4e-3, rel_l2(a1[4], want)


def test_attention_backward_fused_bias_gradients(dev):
    """sam2b200_attn_bwd_ex: the q / k / v bias gradients (column sums of dq / dk / dv over all rows, after the
    conjugate rotation) are accumulated inside the gradient epilogues -- ragged sizes, bf16 and fp32 outputs."""
    from sam2_video_training_b200 import ops
    from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
    g = torch.Generator(device="cuda").manual_seed(9)
    grid, b, nptr = 12, 3, 8                       # N = 144 (not a multiple of 128), M = 2 N + 8
    n, m = grid * grid, 2 * grid * grid + nptr
    table = compute_axial_cis(dim=256, end_x=grid, end_y=grid).to(dev)
    q = (torch.randn(b, n, 256, device=dev, generator=g)).to(torch.bfloat16)
    k = (torch.randn(b, m, 256, device=dev, generator=g)).to(torch.bfloat16)
    v = torch.randn(b, m, 256, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(b, n, 256, device=dev, generator=g).to(torch.bfloat16)
    o, o32, lse = ops.attn_fwd(q, k, v, 1 / 16.0)
    for gdt in (torch.float32, torch.bfloat16):
        db = [torch.full((256,), 0.5, device=dev) for _ in range(3)]      # accumulate on top of existing content
        dq, dk, dv = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=m - nptr, grad_dtype=gdt,
                                  dbias=tuple(db))
        dq0, dk0, dv0 = ops.attn_bwd(q, k, v, None, o32, do, lse, 1 / 16.0, table=table, n_rope_k=m - nptr, grad_dtype=torch.float32)
        for got, ref, name in zip(db, (dq0, dk0, dv0), "qkv"):
            want = ref.double().sum(dim=(0, 1)) + 0.5
            assert rel_l2(got.double(), want) < (2e-3 if gdt == torch.bfloat16 else 1e-5), (name, gdt)
        assert rel_l2(dq.float(), dq0) < 6e-3 and rel_l2(dk.float(), dk0) < 6e-3 and rel_l2(dv.float(), dv0) < 6e-3


@pytest.mark.parametrize("tag", [t for t, _ in detgen.bank_scenarios()])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_memory_bank_assembly_matches_oracle(dev, golden_dir, tag, dtype):
    """sam2b200_bank_gather + the host selection (memory_bank.assemble_memory) against the reference-generated fixture
    (fp32: bit-exact memory, memory_pos to 1 ulp) and the oracle (bf16 frame features; gradients to maskmem_tpos_enc and
    obj_ptr_tpos_proj)."""
    import numpy as np
    from oracle import bank_oracle as bo
    from sam2_video_training_b200 import memory_bank as mb
    kw = dict(detgen.bank_scenarios())[tag]
    od, tpos, pw, pb = detgen.bank_inputs(kw["cond"], kw["non_cond"])

    def cast(o):
        return {"maskmem_features": o["maskmem_features"].to(dtype), "maskmem_pos_enc": [o["maskmem_pos_enc"][0].to(dtype)],
                "obj_ptr": o["obj_ptr"]}
    od = {k_: {t: cast(o) for t, o in v_.items()} for k_, v_ in od.items()}
    od_dev = {k_: {t: {"maskmem_features": o["maskmem_features"].to(dev), "maskmem_pos_enc": [o["maskmem_pos_enc"][0].to(dev)],
                       "obj_ptr": o["obj_ptr"].to(dev)} for t, o in v_.items()} for k_, v_ in od.items()}
    args = dict(max_cond_frames_in_attn=kw.get("max_cond", -1), memory_temporal_stride_for_eval=kw.get("stride", 1))
    proj = torch.nn.Linear(256, 64).to(dev)
    with torch.no_grad():
        proj.weight.copy_(pw); proj.bias.copy_(pb)
    tpos_d = tpos.to(dev).requires_grad_(True)
    memory, memory_pos, n_ptr = mb.assemble_memory(mb.BankConfig(**args), kw["frame_idx"], od_dev, kw["num_frames"], tpos_d, proj,
                                                   training=kw["training"], track_in_reverse=kw.get("reverse", False))
    to, pwo, pbo = (t.clone().requires_grad_(True) for t in (tpos, pw, pb))
    rm, rp, rn = bo.assemble_memory(bo.BankConfig(**args), kw["frame_idx"], od, kw["num_frames"], to, pwo, pbo, kw["training"],
                                    track_in_reverse=kw.get("reverse", False))
    assert n_ptr == rn and memory.shape == rm.shape and memory.dtype == torch.float32
    assert torch.equal(memory.cpu(), rm.float())
    assert torch.allclose(memory_pos.cpu(), rp.float(), atol=2e-6, rtol=0)
    if dtype == torch.float32:
        g = np.load(os.path.join(golden_dir, f"bank_{tag}.npz"))
        assert int(g["n_ptr"]) == n_ptr and torch.equal(memory.cpu(), torch.from_numpy(g["memory"]))
        assert torch.allclose(memory_pos.cpu(), torch.from_numpy(g["memory_pos"]), atol=2e-6, rtol=0)
    w = detgen.det(tuple(rp.shape), 0.013, 0.9, 1.0)
    (memory_pos * w.to(dev)).sum().backward()
    (rp.float() * w).sum().backward()
    assert rel_l2(tpos_d.grad.cpu(), to.grad) < 1e-5
    assert rel_l2(proj.weight.grad.cpu(), pwo.grad) < 1e-5 and rel_l2(proj.bias.grad.cpu(), pbo.grad) < 1e-5


@pytest.mark.parametrize("tag", [t for t, _ in detgen.bank_scenarios()])
def test_packed_bank_matches_reference_fixture(dev, golden_dir, tag):
    """memory_bank.assemble_memory_packed (sam2b200_bank_gather_packed): the bank written directly in the kernel layout --
    memk = bf16(memory + memory_pos), memv = bf16(memory), [B, M, 64] batch-first -- against the fixture produced by the
    UNMODIFIED SAM2Base._prepare_memory_conditioned_features (bit-exact on memv, 1 bf16 ulp on memk)."""
    import numpy as np
    from sam2_video_training_b200 import memory_bank as mb
    kw = dict(detgen.bank_scenarios())[tag]
    od, tpos, pw, pb = detgen.bank_inputs(kw["cond"], kw["non_cond"])
    od_dev = {k_: {t: {"maskmem_features": o["maskmem_features"].to(dev), "maskmem_pos_enc": [o["maskmem_pos_enc"][0].to(dev)],
                       "obj_ptr": o["obj_ptr"].to(dev)} for t, o in v_.items()} for k_, v_ in od.items()}
    args = dict(max_cond_frames_in_attn=kw.get("max_cond", -1), memory_temporal_stride_for_eval=kw.get("stride", 1))
    g = np.load(os.path.join(golden_dir, f"bank_{tag}.npz"))
    ref_mem, ref_pos = torch.from_numpy(g["memory"]), torch.from_numpy(g["memory_pos"])          # [M, B, 64] fp32
    proj = torch.nn.Linear(256, 64).to(dev)
    with torch.no_grad():
        proj.weight.copy_(pw); proj.bias.copy_(pb)
    tpos_d = tpos.to(dev).requires_grad_(True)
    out = mb.assemble_memory_packed(mb.BankConfig(**args), kw["frame_idx"], od_dev, kw["num_frames"], tpos_d, proj, training=kw["training"],
                                    track_in_reverse=kw.get("reverse", False))
    assert isinstance(out, mb.PackedBank) and out.num_obj_ptr_tokens == int(g["n_ptr"]) and out.shape == tuple(ref_mem.shape)
    assert out.memk.dtype == out.memv.dtype == torch.bfloat16 and not out.memv.requires_grad and not out.memk.requires_grad
    assert torch.equal(out.memv.cpu(), ref_mem.transpose(0, 1).to(torch.bfloat16))
    want_k = (ref_mem + ref_pos).transpose(0, 1)
    assert float((out.memk.float().cpu() - want_k).abs().max()) <= 2.0 ** -7 * float(want_k.abs().max())
    # the bank's differentiable part: rows of maskmem_tpos_enc / projected pointer positions (gradients flow through the stack,
    # test_memory_attention_on_packed_bank_matches_reference_layout)
    assert out.tpos_rows.requires_grad and out.tpos_rows.shape == (out.n_slots, 64)
    assert out.n_slots * out.hw + out.num_obj_ptr_tokens == out.memk.shape[1]


def test_memory_attention_on_packed_bank_matches_reference_layout(dev):
    """MemoryAttention fed a PackedBank (no fp32 [M, B, 64] pair, no re-pack inside the stack) == MemoryAttention fed the
    reference-layout memory / memory_pos built from the same frames: output, d curr and the gradient that reaches
    maskmem_tpos_enc through the bank; eager and CUDA-graph replay."""
    from sam2_video_training_b200 import memory_bank as mb
    from sam2_video_training_b200.graphs import GraphedMemoryAttention
    from sam2_video_training_b200.modeling.memory_attention import build_memory_attention
    torch.manual_seed(5)
    g = torch.Generator(device="cuda").manual_seed(9)
    grid, b, nf = 8, 64, 4                          # 64 objects: the raw-memory cross-attention path
    n = grid * grid
    model = build_memory_attention(dropout=0.0).to(dev).train()
    graphed = GraphedMemoryAttention(model)
    od = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
    for t in range(nf):
        o = {"maskmem_features": torch.randn(b, 64, grid, grid, device=dev, generator=g),
             "maskmem_pos_enc": [torch.randn(b, 64, grid, grid, device=dev, generator=g) * 0.7],
             "obj_ptr": torch.randn(b, 256, device=dev, generator=g)}
        od["cond_frame_outputs" if t == 0 else "non_cond_frame_outputs"][t] = o
    proj = torch.nn.Linear(256, 64).to(dev)
    curr0 = torch.randn(n, b, 256, device=dev, generator=g)
    cpos = torch.randn(n, b, 256, device=dev, generator=g) * 0.7
    gout = torch.randn(n, b, 256, device=dev, generator=g)
    res = {}
    for mode in ("reference_layout", "packed", "packed_graph", "packed_graph_replay"):
        model.zero_grad(set_to_none=True)
        proj.zero_grad(set_to_none=True)
        tpos = (torch.randn(7, 1, 1, 64, device=dev, generator=torch.Generator(device="cuda").manual_seed(1)) * 0.02).requires_grad_(True)
        curr = curr0.clone().requires_grad_(True)
        if mode == "reference_layout":
            memory, memory_pos, p = mb.assemble_memory(mb.BankConfig(), nf, od, 10, tpos, proj, training=True)
            out = model(curr, memory.detach(), cpos, memory_pos, p)
        else:
            bank = mb.assemble_memory_packed(mb.BankConfig(), nf, od, 10, tpos, proj, training=True)
            out = (graphed if "graph" in mode else model)(curr, bank, cpos)
        out.backward(gout)
        torch.cuda.synchronize()
        res[mode] = (out.detach().clone(), curr.grad.clone(), tpos.grad.clone(), proj.weight.grad.clone(),
                     torch.cat([p_.grad.flatten() for p_ in model.parameters()]))
    ref = res["reference_layout"]
    for mode in ("packed", "packed_graph", "packed_graph_replay"):
        got = res[mode]
        assert rel_l2(got[0], ref[0]) < 2e-3, (mode, rel_l2(got[0], ref[0]))            # memk rounded once (fp32 add order differs)
        assert cosine(got[1], ref[1]) > 0.9999 and cosine(got[4], ref[4]) > 0.9999, mode
        assert cosine(got[2], ref[2]) > GRAD_COS_TOL and cosine(got[3], ref[3]) > GRAD_COS_TOL, (mode, cosine(got[2], ref[2]))
    assert len(graphed._graphs) == 1


# ---- mask loss fused with its producer side (4x bilinear up-sampling + category merge), SURVEY.md section 8f rank 2 ----

MERGED_W = {"loss_mask": 20, "loss_dice": 1, "loss_iou": 1, "loss_class": 0}
MERGED_MODES = (("l1", dict(iou_use_l1_loss=True)), ("mse", dict(iou_use_l1_loss=False)),
                ("temp", dict(iou_use_l1_loss=True, logit_temperature=1.6, focal_alpha=0.6)))


def _merged_stages(low, ious):
    return [{"multistep_pred_multimasks": [low[f]], "multistep_pred_ious": [ious[f]]} for f in range(len(low))]


@pytest.mark.parametrize("tag", ["t2_n5_c4_s6", "t2_n7_c3_s11"])
def test_merged_loss_vs_reference_golden(dev, golden_dir, tag):
    """Fused up-sample + merge + loss against fixtures produced by the unmodified reference chain (F.interpolate,
    merge_object_results_to_category, MultiStepMultiMasksAndIous): values <= 1e-3 relative (measured ~1e-6),
    gradients w.r.t. the low-res logits and the per-object IoU predictions."""
    from sam2_video_training_b200.merged_loss import CategoryMergedMultiStepLoss
    g = np.load(os.path.join(golden_dir, f"merged_{tag}.npz"))
    t, n, c, s = int(g["t"]), int(g["n_obj"]), int(g["c"]), int(g["s"])
    low, ip, o2c, tg = detgen.merged_inputs(t, n, c, s)
    for mode, kw in MERGED_MODES:
        crit = CategoryMergedMultiStepLoss(dict(MERGED_W), supervise_all_iou=True, **kw)
        x = [low[f].to(dev).requires_grad_(True) for f in range(t)]
        p = [ip[f].to(dev).requires_grad_(True) for f in range(t)]
        losses = crit(_merged_stages(x, p), o2c, c, tg.to(dev))
        losses["total_loss"].backward()
        for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
            ref = float(g[f"{mode}:{k}"])
            assert abs(float(losses[k]) - ref) <= 1e-5 * max(1.0, abs(ref)), (mode, k)
        dlow = torch.stack([v.grad for v in x]).cpu().numpy()
        diou = torch.stack([v.grad for v in p]).cpu().numpy()
        assert rel_l2(torch.from_numpy(dlow), torch.from_numpy(g[f"{mode}:dlow"])) < 2e-5, mode
        np.testing.assert_allclose(diou, g[f"{mode}:diou"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("t,n_obj,c,s,o2c", [
    (3, 6, 4, 96, [0, 2, 0, 1, 2, 2]),          # cfg2 resolution (384 px), category 3 without objects
    (2, 3, 3, 50, [2, 0, 1]),                    # ragged tiles (s not a multiple of 8 / 32), one object per category
    (1, 4, 1, 5, [0, 0, 0, 0]),                  # smaller than one tile, all objects in one category
    (1, 9, 2, 128, [1, 1, 1, 1, 1, 1, 1, 1, 0]),  # 512 px, 8 objects in one category
])
def test_merged_loss_vs_oracle_random(dev, t, n_obj, c, s, o2c):
    from oracle import merge_oracle as mo
    from sam2_video_training_b200.merged_loss import CategoryMergedMultiStepLoss
    g = torch.Generator().manual_seed(100 + s)
    low = torch.randn(t, n_obj, 1, s, s, generator=g) * 4
    ip = torch.rand(t, n_obj, 1, generator=g)
    S = 4 * s
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    tg = torch.zeros(t, c, S, S, dtype=torch.bool)
    for f in range(t):
        for ch in range(c):
            if c > 2 and f == 1 and ch == 1:
                continue                          # an empty channel: exercises the valid filter
            cx, cy = S * (0.3 + 0.4 * torch.rand((), generator=g)), S * (0.3 + 0.4 * torch.rand((), generator=g))
            r = S * (0.1 + 0.2 * torch.rand((), generator=g))
            tg[f, ch] = (xx - cx) ** 2 + (yy - cy) ** 2 < r * r
    kw = dict(iou_use_l1_loss=False)
    xo = low.clone().requires_grad_(True)
    po = ip.clone().requires_grad_(True)
    ref = mo.merged_multistep_loss([xo[f] for f in range(t)], [po[f] for f in range(t)], o2c, c, tg, dict(MERGED_W), **kw)
    ref["total_loss"].backward()
    crit = CategoryMergedMultiStepLoss(dict(MERGED_W), supervise_all_iou=True, **kw)
    x = [low[f].to(dev).requires_grad_(True) for f in range(t)]
    p = [ip[f].to(dev).requires_grad_(True) for f in range(t)]
    out = crit(_merged_stages(x, p), o2c, c, tg.to(dev))
    out["total_loss"].backward()
    for k in ("loss_mask", "loss_dice", "loss_iou", "total_loss"):
        assert abs(float(out[k]) - float(ref[k])) <= 1e-4 * max(1.0, abs(float(ref[k]))), k   # bound from BASELINE.json: 1e-3
    dlow = torch.stack([v.grad for v in x]).cpu()
    assert rel_l2(dlow, xo.grad) < 1e-4
    cosv = torch.nn.functional.cosine_similarity(dlow.flatten().double(), xo.grad.flatten().double(), dim=0)
    assert cosv > 0.99999
    np.testing.assert_allclose(torch.stack([v.grad for v in p]).cpu().numpy(), po.grad.numpy(), rtol=1e-3, atol=1e-6)
    # the merged op equals the un-fused pipeline of this package: up-sample + max on the GPU with torch, then the fused loss
    from sam2_video_training_b200.losses import MultiStepMultiMasksAndIous
    hi = [torch.nn.functional.interpolate(low[f].to(dev), size=(S, S), mode="bilinear", align_corners=False) for f in range(t)]
    groups = mo.category_groups(o2c, c)
    merged = [mo.merge_frame(hi[f], ip[f].to(dev), groups) for f in range(t)]
    outs = [{"multistep_pred_multimasks_high_res": [m[0]], "multistep_pred_ious": [m[1]],
             "multistep_object_score_logits": [None]} for m in merged]
    un = MultiStepMultiMasksAndIous(dict(MERGED_W), supervise_all_iou=True, **kw)(outs, tg.to(dev))
    assert abs(float(un["total_loss"]) - float(out["total_loss"])) <= 1e-5 * abs(float(un["total_loss"]))


def test_merged_loss_error_contract(dev):
    from sam2_video_training_b200.merged_loss import CategoryMergedMultiStepLoss
    with pytest.raises(NotImplementedError):
        CategoryMergedMultiStepLoss(dict(MERGED_W), pred_obj_scores=True)
    with pytest.raises(ValueError):
        CategoryMergedMultiStepLoss(dict(MERGED_W), logit_temperature=0)
    crit = CategoryMergedMultiStepLoss(dict(MERGED_W))
    low = torch.randn(2, 1, 8, 8, device=dev)
    iou = torch.rand(2, 1, device=dev)
    tg = torch.zeros(1, 2, 32, 32, dtype=torch.bool, device=dev)
    with pytest.raises(ValueError, match="No valid masks"):           # losses.py:161
        crit(_merged_stages([low], [iou]), [0, 1], 2, tg)
    with pytest.raises(ValueError):                                    # wrong target resolution
        crit(_merged_stages([low], [iou]), [0, 1], 2, torch.ones(1, 2, 16, 16, dtype=torch.bool, device=dev))
    with pytest.raises(IndexError):
        crit(_merged_stages([low], [iou]), [0, 5], 2, torch.ones(1, 2, 32, 32, dtype=torch.bool, device=dev))
    with pytest.raises(AssertionError):                                # losses.py:113
        crit(_merged_stages([low, low], [iou, iou]), [0, 1], 2, tg)


# ---- memory encoder -------------------------------------------------------------------------

@pytest.mark.parametrize("c,p_,act", [(4, 1000, True), (16, 777, True), (64, 4097, True), (256, 513, True), (256, 64, False)])
def test_ln_gelu_kernels(dev, c, p_, act):
    """sam2b200_ln_gelu_fwd / _bwd (LayerNorm2d + exact GELU over the channels of channels-last pixels) against fp64 torch."""
    from sam2_video_training_b200.modeling.memory_encoder import _LnGeluFn
    g = torch.Generator(device="cuda").manual_seed(c + p_)
    x = (torch.randn(p_, c, device=dev, generator=g) * 1.7 + 0.3).requires_grad_(True)
    w = (1 + 0.2 * torch.randn(c, device=dev, generator=g)).requires_grad_(True)
    b = (0.2 * torch.randn(c, device=dev, generator=g)).requires_grad_(True)
    go = torch.randn(p_, c, device=dev, generator=g)
    y = _LnGeluFn.apply(x, w, b, 1e-6, act)
    y.backward(go)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    z = torch.nn.functional.layer_norm(xd, (c,), wd, bd, 1e-6)
    ref = torch.nn.functional.gelu(z) if act else z
    ref.backward(go.double())
    assert rel_l2(y, ref) < 2e-6
    assert rel_l2(x.grad, xd.grad) < 2e-5 and rel_l2(w.grad, wd.grad) < 2e-5 and rel_l2(b.grad, bd.grad) < 2e-5


@pytest.mark.parametrize("b,h,w", [(2, 24, 24), (3, 5, 9), (1, 32, 32)])
def test_dwconv7_kernels(dev, b, h, w):
    """sam2b200_dwconv7 / _bwd_w (depth-wise 7 x 7, padding 3, channels-last) against F.conv2d(groups = C) in fp64."""
    from sam2_video_training_b200.modeling.memory_encoder import _DwConv7Fn
    g = torch.Generator(device="cuda").manual_seed(b * h + w)
    c = 256
    x = torch.randn(b, h, w, c, device=dev, generator=g).requires_grad_(True)
    wt = (torch.randn(c, 1, 7, 7, device=dev, generator=g) / 7).requires_grad_(True)
    bias = (0.1 * torch.randn(c, device=dev, generator=g)).requires_grad_(True)
    go = torch.randn(b, h, w, c, device=dev, generator=g)
    y = _DwConv7Fn.apply(x, wt, bias)
    y.backward(go)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, wt, bias))
    ref = torch.nn.functional.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=3, groups=c).permute(0, 2, 3, 1)
    ref.backward(go.double())
    assert rel_l2(y, ref) < 2e-6
    assert rel_l2(x.grad, xd.grad) < 2e-6 and rel_l2(wt.grad, wd.grad) < 2e-5 and rel_l2(bias.grad, bd.grad) < 2e-5


@pytest.mark.parametrize("tag", ["b2_g4_sigmoid", "b3_g6_scaled"])
def test_memory_encoder_vs_reference_golden(dev, golden_dir, tag):
    """MemoryEncoder (B200 path) against the UNMODIFIED reference class (tests/golden/memenc_*.npz): same state_dict keys, the
    reference's own init reproduced by constructing the module under torch.manual_seed(0); features within 1e-2 relative
    (bf16 tensor-core GEMMs vs the reference's fp32), position encoding exact, input / parameter gradient cosines >= 0.999."""
    from oracle import memenc_oracle as mo
    from sam2_video_training_b200.modeling.memory_encoder import build_memory_encoder
    g = np.load(os.path.join(golden_dir, f"memenc_{tag}.npz"))
    b, grid, seed, skip = int(g["b"]), int(g["grid"]), int(g["seed"]), bool(int(g["skip"]))
    torch.manual_seed(0)
    model = build_memory_encoder()
    names = [str(n) for n in g["param_names"]]
    assert [n for n, _ in model.named_parameters()] == names          # the reference's state_dict layout
    sd = mo.reference_init_state(0)
    with torch.no_grad():
        for i in range(2):
            model.fuser.layers[i].gamma.copy_(sd[f"fuser.layers.{i}.gamma"])
    for n, p in model.named_parameters():                              # same RNG consumption as the reference constructors
        assert torch.equal(p.detach(), sd[n]), n
    model = model.to(dev).train()
    inp = mo.random_inputs(b, grid, seed)
    pix = inp["pix_feat"].to(dev).requires_grad_(True)
    masks = inp["masks"].to(dev).requires_grad_(True)
    m_in = torch.sigmoid(masks) * 20.0 - 10.0 if skip else masks
    out = model(pix, m_in, skip_mask_sigmoid=skip)
    feat, pos = out["vision_features"], out["vision_pos_enc"][0]
    feat.backward(inp["grad_out"].to(dev))
    torch.cuda.synchronize()
    assert feat.shape == (b, 64, grid, grid) and pos.shape == feat.shape
    assert rel_l2(feat, torch.from_numpy(g["features"])) < ATTN_REL_TOL, rel_l2(feat, torch.from_numpy(g["features"]))
    assert rel_l2(pos, torch.from_numpy(g["pos"])) < 1e-6
    assert cosine(pix.grad, torch.from_numpy(g["d_pix_feat"])) > GRAD_COS_TOL
    assert cosine(masks.grad, torch.from_numpy(g["d_masks"])) > GRAD_COS_TOL
    named = dict(model.named_parameters())
    for key in g.files:
        if key.startswith("dparam:"):
            c_ = cosine(named[key[7:]].grad, torch.from_numpy(g[key]))
            assert c_ > GRAD_COS_TOL, (key, c_)
    for n, s_ in zip(names, g["param_grad_abs_sums"]):
        mine = float(named[n].grad.double().abs().sum())
        assert abs(mine - s_) <= 5e-2 * max(abs(s_), 1e-3), (n, mine, s_)


def test_memory_encoder_cfg2_shape_vs_oracle(dev):
    """cfg2 frame shape (7 objects, 384 px masks -> 24 x 24 memory features) against oracle/memenc_oracle.py run in fp32 on the
    device (TF32 off): the features feed the memory bank, so the output tolerance is the attention path's 1e-2."""
    from oracle import memenc_oracle as mo
    from sam2_video_training_b200.modeling.memory_encoder import build_memory_encoder
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = build_memory_encoder()
    sd = mo.reference_init_state(0)
    with torch.no_grad():
        for i in range(2):
            model.fuser.layers[i].gamma.copy_(sd[f"fuser.layers.{i}.gamma"])
    model = model.to(dev).train()
    inp = {k: v.to(dev) for k, v in mo.random_inputs(7, 24, 5).items()}
    pix = inp["pix_feat"].clone().requires_grad_(True)
    masks = inp["masks"].clone().requires_grad_(True)
    out = model(pix, torch.sigmoid(masks) * 20.0 - 10.0, skip_mask_sigmoid=True)
    out["vision_features"].backward(inp["grad_out"])
    p = {k: v.to(dev).clone().requires_grad_(True) for k, v in sd.items()}
    pix_o = inp["pix_feat"].clone().requires_grad_(True)
    masks_o = inp["masks"].clone().requires_grad_(True)
    feat, pos = mo.memory_encoder(p, pix_o, torch.sigmoid(masks_o) * 20.0 - 10.0, skip_mask_sigmoid=True)
    feat.backward(inp["grad_out"])
    torch.cuda.synchronize()
    assert rel_l2(out["vision_features"], feat) < ATTN_REL_TOL and rel_l2(out["vision_pos_enc"][0], pos) < 1e-6
    assert cosine(pix.grad, pix_o.grad) > GRAD_COS_TOL and cosine(masks.grad, masks_o.grad) > GRAD_COS_TOL
    mine = torch.cat([q.grad.flatten() for _, q in model.named_parameters()])
    theirs = torch.cat([p[n].grad.flatten() for n, _ in model.named_parameters()])
    assert cosine(mine, theirs) > GRAD_COS_TOL, cosine(mine, theirs)
