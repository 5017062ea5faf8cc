"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/sam2_b200.h
declares; the host-side mirror keeps the reference's interface; no silent CPU fallback."""
import os
import re

import pytest
import torch

import sam2_video_training_b200 as pkg
from sam2_video_training_b200 import _lib
from sam2_video_training_b200 import build as pkg_build
from sam2_video_training_b200.losses import BCECategoryLoss, CORE_LOSS_KEY, MultiStepMultiMasksAndIous
from sam2_video_training_b200.modeling.memory_attention import MemoryAttention, build_memory_attention
from sam2_video_training_b200.modeling.position_encoding import compute_axial_cis
from sam2_video_training_b200.modeling.sam.transformer import Attention, RoPEAttention

from oracle import attention_oracle as ao
from oracle import detgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    pkg_build.build(verbose=False)
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sam2_b200.h")).read()
    declared = set(re.findall(r"\b(sam2b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sam2b200_version() >= 100


def test_no_gpu_compute_without_device(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.sam2b200_check_device(0) != 0
    assert lib.sam2b200_last_error()


def test_state_dict_matches_reference_layout():
    m = build_memory_attention()
    sd = m.state_dict()
    assert len(sd) == 106
    assert sum(p.numel() for p in m.parameters()) == 5922304
    names = [n for n, _ in m.named_parameters()]
    assert names == [n for n, _ in detgen.param_shapes()]
    for n, shape in detgen.param_shapes():
        assert tuple(sd[n].shape) == tuple(shape), n
    assert "freqs_cis" not in "".join(sd.keys())  # plain attribute, transformer.py:269-272
    assert sd["layers.0.cross_attn_image.k_proj.weight"].shape == (256, 64)
    assert isinstance(m.layers[0].cross_attn_image, RoPEAttention)
    assert isinstance(m.layers[0].self_attn, Attention)


def test_rope_table_matches_oracle():
    for n in (16, 576, 1024):
        w = int(n ** 0.5)
        t = compute_axial_cis(dim=256, end_x=w, end_y=w)
        cos, sin = ao.axial_rope_table(n)
        assert t.shape == (n, 128, 2)
        assert torch.equal(t[..., 0], cos) and torch.equal(t[..., 1], sin)


def test_cpu_tensors_fail_loudly(lib):
    m = build_memory_attention().eval()
    inp = detgen.attention_inputs(4, 1, 1, 4)
    with pytest.raises(_lib.Sam2B200Error):
        m(inp["curr"], inp["memory"], inp["curr_pos"], inp["memory_pos"], 4)
    crit = MultiStepMultiMasksAndIous({"loss_mask": 20, "loss_dice": 1, "loss_iou": 1})
    logits, targets, iou = detgen.loss_inputs(1, 2, 8)
    outs = [{"multistep_pred_multimasks_high_res": [logits[0]], "multistep_pred_ious": [iou[0]],
             "multistep_object_score_logits": [torch.zeros(2, 1)]}]
    with pytest.raises(_lib.Sam2B200Error):
        crit(outs, targets)


def test_reference_error_contract():
    with pytest.raises(ValueError):
        MultiStepMultiMasksAndIous({"loss_mask": 1, "loss_dice": 1, "loss_iou": 1}, logit_temperature=0)
    with pytest.raises(ValueError):
        BCECategoryLoss(logit_temperature=-1.0)
    with pytest.raises(AssertionError):
        MultiStepMultiMasksAndIous({"loss_mask": 1, "loss_dice": 1})
    crit = MultiStepMultiMasksAndIous({"loss_mask": 1, "loss_dice": 1, "loss_iou": 1})
    assert crit.weight_dict["loss_class"] == 0.0
    assert CORE_LOSS_KEY == "total_loss"
    m = build_memory_attention()
    with pytest.raises(AssertionError):  # batch mismatch, memory_attention.py:135-137
        m(torch.zeros(16, 2, 256), torch.zeros(20, 3, 64), torch.zeros(16, 2, 256), torch.zeros(20, 3, 64))
    with pytest.raises(_lib.Sam2B200Error):
        Attention(256, 8)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.Sam2B200Error):
        _lib.load()


def test_product_never_imports_oracle():
    pkg_dir = os.path.dirname(pkg.__file__)
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_argument_validation_needs_no_device(lib):
    """Entry points reject bad arguments before touching the device: error code + message, no exception, no launch."""
    assert lib.sam2b200_merged_loss_workspace_bytes(2, 3, 4, 16) > 0
    assert lib.sam2b200_merged_loss_workspace_bytes(0, 3, 4, 16) == 0
    rc = lib.sam2b200_merged_loss_fwd(None, None, None, None, None, None, None, None, None, None, None, None,
                                      2, 3, 4, 16, 0.25, 2.0, 1.0, 1, None)
    assert rc == -1 and b"merged_loss_fwd" in lib.sam2b200_last_error()
    rc = lib.sam2b200_merged_loss_bwd(None, None, None, None, None, None, None, None, None, None, None, None, None,
                                      2, 3, 4, 16, 0.25, 2.0, 1.0, 1, None)
    assert rc == -1 and b"merged_loss_bwd" in lib.sam2b200_last_error()
    rc = lib.sam2b200_mask_loss_fwd(None, None, None, None, None, None, None, None, 1, 1, 16, 0, 0.25, 2.0, 1.0, 0, 1, None)
    assert rc == -1 and b"mask_loss_fwd" in lib.sam2b200_last_error()


def test_merged_loss_host_checks_without_device():
    """Host-side contract of CategoryMergedMultiStepLoss that needs no GPU: constructor checks, CPU tensors fail loudly."""
    from sam2_video_training_b200 import _lib as L
    from sam2_video_training_b200.merged_loss import CategoryMergedMultiStepLoss
    w = {"loss_mask": 20, "loss_dice": 1, "loss_iou": 1}
    with pytest.raises(NotImplementedError):
        CategoryMergedMultiStepLoss(dict(w), pred_obj_scores=True)
    with pytest.raises(ValueError):
        CategoryMergedMultiStepLoss(dict(w), logit_temperature=-1.0)
    with pytest.raises(AssertionError):
        CategoryMergedMultiStepLoss({"loss_mask": 1})
    crit = CategoryMergedMultiStepLoss(dict(w))
    assert crit.weight_dict["loss_class"] == 0.0
    stages = [{"multistep_pred_multimasks": [torch.zeros(2, 1, 4, 4)], "multistep_pred_ious": [torch.zeros(2, 1)]}]
    with pytest.raises(L.Sam2B200Error):
        crit(stages, [0, 1], 2, torch.ones(1, 2, 16, 16, dtype=torch.bool))


def test_single_flag_switch_against_the_real_reference_module():
    """integrate.use_b200_attention on a stand-in for SAM2Base that holds the UNMODIFIED reference MemoryAttention (built
    through oracle/ref_shim.py; skipped where /root/reference does not exist): strict state_dict transfer both ways, the
    freeze map and train/eval mode are kept, the attributes the caller touches exist (memory_attention.py:119-169)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference sources not present")
    from sam2_video_training_b200 import integrate
    from sam2_video_training_b200.modeling.memory_attention import MemoryAttention
    from sam2_video_training_b200.modeling.sam.transformer import RoPEAttention as FastRope

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.memory_attention = ref_shim.build_memory_attention(dropout=0.1)

    m = Holder().eval()
    for p in m.memory_attention.layers[1].parameters():
        p.requires_grad_(False)
    ref_sd = {k: v.clone() for k, v in m.memory_attention.state_dict().items()}
    ref_mod = m.memory_attention
    fast = integrate.use_b200_attention(m)
    assert m.memory_attention is fast and isinstance(fast, MemoryAttention) and not fast.training
    assert list(fast.state_dict().keys()) == list(ref_sd.keys())
    for k, v in fast.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    assert [p.requires_grad for p in fast.parameters()] == [p.requires_grad for p in ref_mod.parameters()]
    assert fast.d_model == 256 and fast.num_layers == 4 and len(fast.layers) == 4 and fast.norm.normalized_shape == (256,)
    assert isinstance(fast.layers[0].cross_attn_image, FastRope) and fast.layers[0].cross_attn_image.rope_k_repeat
    assert fast.layers[0].dropout1.p == pytest.approx(0.1) if hasattr(fast.layers[0], "dropout1") else True
    ref_mod.load_state_dict(fast.state_dict(), strict=True)          # and back: checkpoints stay interchangeable
    with pytest.raises(ValueError):
        integrate.b200_criterion("nope")
    crit = integrate.b200_criterion("multi_step_b200", weight_dict={"loss_mask": 20, "loss_dice": 1, "loss_iou": 1})
    assert crit.weight_dict["loss_class"] == 0.0


def test_bench_reference_arm_runs_on_cpu_and_keeps_the_json_contract():
    """`bench.py --impl reference` (the CPU oracle timed on the host cores) needs no GPU and prints ONE JSON line with the
    contract keys; under torchrun ranks other than 0 exit 0 without work."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    # cfg1 (one object) keeps the CPU suite short; the default workload is checked through the argument parser below
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "cfg1_384px_T10_1obj_x1clip"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "memory_attn_fwd_bwd_plus_mask_loss_clip_frames_per_sec"
    assert d["unit"] == "clip-frames/s" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    # the unmodified reference modules when baseline/_ref is installed (scripts/install_reference.py), else the oracle port
    has_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "sam2_video", "model", "modeling", "memory_attention.py"))
    assert d["config"]["workload"].startswith("cfg1") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] == ("reference" if has_ref else "port")
    import bench
    assert bench.WORKLOADS and list(bench.WORKLOADS)[0].startswith("cfg2")     # default workload = BASELINE configs[1]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    env1 = dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env1, cwd=ROOT)
    assert out1.returncode == 0 and not [l for l in out1.stdout.splitlines() if l.startswith("{")]


def test_memory_encoder_drop_in_loads_the_reference_state_dict():
    """integrate.use_b200_memory_encoder: the drop-in takes the reference MemoryEncoder's 40 tensors with strict=True and keeps
    the freeze map; without a GPU the forward refuses to run (no CPU fallback)."""
    import torch
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference not present")
    from sam2_video_training_b200 import _lib
    from sam2_video_training_b200.integrate import use_b200_memory_encoder
    torch.manual_seed(0)
    holder = torch.nn.Module()
    holder.memory_encoder = ref_shim.build_memory_encoder()
    ref_sd = {k: v.clone() for k, v in holder.memory_encoder.state_dict().items()}
    holder.memory_encoder.fuser.layers[1].gamma.requires_grad_(False)
    fast = use_b200_memory_encoder(holder)
    assert holder.memory_encoder is fast and list(fast.state_dict().keys()) == list(ref_sd.keys()) and len(ref_sd) == 40
    for k, v in fast.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    assert not fast.fuser.layers[1].gamma.requires_grad and fast.fuser.layers[0].gamma.requires_grad
    if not torch.cuda.is_available():
        with pytest.raises(_lib.Sam2B200Error):
            fast(torch.zeros(1, 256, 2, 2), torch.zeros(1, 1, 32, 32))


def test_gemm_argument_validation_needs_no_device(lib):
    """sam2b200_gemm / gemm_ex reject malformed problems on the host (no launch, no device): widths other than 64 / k x 256,
    K not a multiple of 64, misaligned pointers, a rotation table without its geometry, an unpaired row-dot argument."""
    ok = dict(c=4096, ldc=256, a=8192, lda=256, b=12288, ldb=256, layout=0, R=128, K=256, No=256)

    def call(**kw):
        a = {**ok, **kw}
        return lib.sam2b200_gemm(a["c"], a["ldc"], a["a"], a["lda"], a["b"], a["ldb"], a["layout"], a["R"], a["K"], a["No"],
                                 kw.get("bias"), kw.get("table"), kw.get("rpi", 1), kw.get("nrope", 0), kw.get("period", 1),
                                 kw.get("dot_rows"), kw.get("dot_out"), None)
    for bad in (dict(No=192), dict(K=100), dict(a=8200), dict(lda=128), dict(layout=2), dict(R=0), dict(table=4096, rpi=0),
                dict(dot_rows=4096), dict(No=64, ldc=64, dot_out=4096)):
        assert call(**bad) == -1, bad
        assert b"gemm" in lib.sam2b200_last_error()
    rc = lib.sam2b200_gemm_ex(4096, None, None, 256, 256, 8192, 256, 12288, 256, 0, 128, 256, 768, None, 0, None, 1, 0, 1, 0, 0.0, None, 0,
                              None, None, None)
    assert rc == -1          # three outputs of width 256 announced (Nout = 768), two of them missing


def test_segment_indicator_reproduces_the_dense_bank_position_gradient():
    """The packed bank's position gradients: S_l = dk_l^T Ind summed per segment, then sum_l S_l^T Wk_l, equals the reference
    route (dense d memk = sum_l dk_l Wk_l, reduced over objects and over the tokens of every slot / pointer) -- host algebra,
    checked on CPU in fp64."""
    from sam2_video_training_b200 import fused_stack as fs
    torch.manual_seed(0)
    b, ns, hw, n_ptr, per, nl = 3, 2, 5, 3, 4, 2
    m = ns * hw + n_ptr * per
    ind = fs.segment_indicator(torch.device("cpu"), b, m, ns, hw, n_ptr).double()
    assert ind.shape == (b * m, 64) and torch.equal(ind.sum(1), torch.ones(b * m, dtype=torch.float64))
    dk = [torch.randn(b * m, 256, dtype=torch.float64) for _ in range(nl)]
    wk = [torch.randn(256, 64, dtype=torch.float64) for _ in range(nl)]
    dense = sum(d @ w for d, w in zip(dk, wk)).view(b, m, 64)
    want_t = dense[:, :ns * hw].reshape(b, ns, hw, 64).sum((0, 2))
    want_p = dense[:, ns * hw:].reshape(b, n_ptr, per, 64).sum((0, 2))
    seg = torch.cat([d.t() @ ind for d in dk], 0)                     # [nl * 256, 64]
    got = seg.t() @ torch.cat(wk, 0)                                   # [64 segments, 64]
    assert torch.allclose(got[:ns], want_t, atol=1e-9) and torch.allclose(got[ns:ns + n_ptr], want_p, atol=1e-9)
    assert float(got[ns + n_ptr:].abs().max()) == 0.0


def test_gemm_plan_fits_every_shape_of_the_stack_and_beyond(lib):
    """sam2b200_gemm_plan (host only): for every (rows, K, Nout, rotation) the stack produces at BASELINE.json's configurations -- and a
    sweep around them -- the chosen variant's shared-memory layout fits one CTA (227 KB), has at least two ring slots, and the grid
    never exceeds the SM count; resident weights exactly when K <= 256; 128-column blocks exactly for rotated outputs."""
    import ctypes
    out = (ctypes.c_longlong * 8)()
    rows = [1, 100, 576, 56 * 576, 13 * 1024, 4 * 4096, 56 * 4060, 4 * 28736, 1 << 20]
    seen = set()
    for r in rows:
        for k in (64, 128, 192, 256, 768, 2048, 4096):
            for nout, rope, period in ((64, 0, 1), (256, 0, 1), (256, 256, 576), (256, 256, 4096), (256, 256, 1000), (768, 512, 576),
                                       (768, 512, 1024), (2048, 0, 1), (1024, 0, 1)):
                rc = lib.sam2b200_gemm_plan(r, k, nout, rope, period, 148, out)
                assert rc == 0, (r, k, nout, rope, lib.sam2b200_last_error())
                bn, mt, eg, slots, bufs, grid, wres, smem = list(out)
                assert smem <= 227 * 1024 and slots >= 2 and 1 <= grid <= 148 and bufs in (1, 2)
                assert wres == (1 if k <= 256 else 0)
                assert bn == (64 if nout == 64 else (128 if rope else 256))
                assert mt == (2 if (bn == 256 and not wres and (r + 127) // 128 > 148) else 1)
                assert eg == (2 if (bn == 128 and wres) else 1)
                if wres:
                    assert grid % (nout // bn) == 0          # a CTA keeps one column block
                seen.add((bn, mt, eg, wres))
    assert len(seen) >= 6
    assert lib.sam2b200_gemm_plan(128, 100, 256, 0, 1, 148, out) == -1 and lib.sam2b200_gemm_plan(128, 256, 192, 0, 1, 148, out) == -1
