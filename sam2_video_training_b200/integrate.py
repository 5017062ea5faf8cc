"""Single-flag switch for the reference wrapper (INTEGRATION.md sections 2-3).

``use_b200_attention(model)`` is what ``SAM2Model.__init__`` would call right after adopting the built SAM2 model's
sub-modules (sam2_video/model/sam2model.py:80-105) when ``model.use_b200_attention: true``: it replaces
``model.memory_attention`` (the ``nn.Module`` attribute set at sam2_video/model/modeling/sam2_base.py:125 and invoked at
:695-709) by the drop-in, carrying the weights over with ``load_state_dict(strict=True)`` -- the drop-in keeps the
reference's 106 state_dict keys, so checkpoints written afterwards load into the reference classes as well.
``b200_criterion(loss_type, ...)`` mirrors the ``loss.type`` dispatch of sam2_video/training/trainer.py:67-94.
"""
from __future__ import annotations

from typing import Any

from torch import nn

from .losses import BCECategoryLoss, MultiStepMultiMasksAndIous
from .merged_loss import CategoryMergedMultiStepLoss
from .modeling.memory_attention import MemoryAttention, build_memory_attention


def _rope_theta(attn) -> float:
    """The reference keeps theta only inside ``compute_cis = partial(compute_axial_cis, dim=..., theta=...)``
    (transformer.py:264-266)."""
    if hasattr(attn, "rope_theta"):
        return float(attn.rope_theta)
    kw = getattr(getattr(attn, "compute_cis", None), "keywords", None) or {}
    return float(kw.get("theta", 10000.0))


def _feat_sizes(attn):
    n = int(getattr(attn, "freqs_cis").shape[0]) if hasattr(attn, "freqs_cis") else 4096
    w = int(round(n ** 0.5))
    return (w, w) if w * w == n else (64, 64)


def use_b200_attention(model: nn.Module, attr: str = "memory_attention") -> MemoryAttention:
    """Swap ``getattr(model, attr)`` (a reference ``MemoryAttention``) for the B200 drop-in, weights included."""
    ref = getattr(model, attr)
    layer0 = ref.layers[0]
    sa, ca = layer0.self_attn, layer0.cross_attn_image
    # every hyper-parameter is read from the reference module (nothing is assumed to be the shipped yaml's value)
    fast = build_memory_attention(
        dropout=float(getattr(layer0, "dropout_value", layer0.dropout1.p if hasattr(layer0, "dropout1") else 0.0)),
        sa_dropout=float(getattr(sa, "dropout_p", 0.0)), ca_dropout=float(getattr(ca, "dropout_p", 0.0)),
        num_layers=int(getattr(ref, "num_layers", len(ref.layers))),
        dim_feedforward=int(getattr(layer0, "dim_feedforward", layer0.linear1.out_features)),
        rope_theta=_rope_theta(sa), feat_sizes=_feat_sizes(sa), ca_rope_theta=_rope_theta(ca),
        pos_enc_at_input=bool(getattr(ref, "pos_enc_at_input", True)))
    for l_new, l_old in zip(fast.layers, ref.layers):
        for nm in ("norm1", "norm2", "norm3"):
            getattr(l_new, nm).eps = getattr(l_old, nm).eps
    fast.norm.eps = ref.norm.eps
    p0 = next(ref.parameters())
    fast = fast.to(device=p0.device)
    missing, unexpected = fast.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    fast.train(ref.training)
    for p_new, p_old in zip(fast.parameters(), ref.parameters()):
        p_new.requires_grad_(p_old.requires_grad)          # keep the freeze map (trainable_modules, sam2model.py:133-137)
    setattr(model, attr, fast)
    return fast


def use_b200_memory_encoder(model: nn.Module, attr: str = "memory_encoder"):
    """Swap ``getattr(model, attr)`` (a reference ``MemoryEncoder``, built at sam2_base.py:126 from
    configs/sam2/sam2.1_hiera_t.yaml:62-85) for the B200 drop-in, weights included (the reference's 40 state_dict keys)."""
    from .modeling.memory_encoder import build_memory_encoder
    ref = getattr(model, attr)
    out_dim = ref.out_proj.out_channels if isinstance(ref.out_proj, nn.Conv2d) else ref.pix_feat_proj.out_channels
    fast = build_memory_encoder(out_dim=out_dim)
    p0 = next(ref.parameters())
    fast = fast.to(device=p0.device)
    missing, unexpected = fast.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    fast.train(ref.training)
    for p_new, p_old in zip(fast.parameters(), ref.parameters()):
        p_new.requires_grad_(p_old.requires_grad)
    setattr(model, attr, fast)
    return fast


def b200_criterion(loss_type: str, **cfg: Any) -> nn.Module:
    """``multi_step_b200`` / ``bce_b200`` / ``multi_step_merged_b200`` (the last one takes the un-merged tracker stages,
    INTEGRATION.md section 3b)."""
    if loss_type == "multi_step_b200":
        return MultiStepMultiMasksAndIous(**cfg)
    if loss_type == "bce_b200":
        return BCECategoryLoss(**cfg)
    if loss_type == "multi_step_merged_b200":
        return CategoryMergedMultiStepLoss(**cfg)
    raise ValueError(f"unknown loss type {loss_type!r}")
