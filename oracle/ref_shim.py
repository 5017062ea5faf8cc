"""Import shim for the *real* reference modules (TEST INFRASTRUCTURE ONLY).

The reference's hot-path files under ``/root/reference/sam2_video/model`` import
``sam2.modeling.*`` (the un-installed pip ``sam2`` package; see
sam2_video/model/modeling/memory_attention.py:12-14 and
sam2_video/model/modeling/sam/transformer.py:15-16).  This shim loads the vendored
files by path and registers them under the ``sam2.modeling.*`` names they expect,
so that the unmodified reference code can be executed on CPU *in the build
container* to (a) validate ``oracle/`` and (b) generate ``tests/golden/*.npz``.

``/root/reference`` does not exist on the GPU box, so nothing that runs there
(``-m gpu`` tests, ``smoke()``, ``bench.py``) may import this module; callers must
check :func:`available` first.  No reference source is copied into this repo.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SAM2_REFERENCE_ROOT", "/root/reference")
_MODELING = os.path.join(REF_ROOT, "sam2_video", "model", "modeling")
_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(_MODELING, "memory_attention.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return a namespace with the reference's hot-path symbols."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not present under {REF_ROOT}")
    # Package skeleton for `sam2`, `sam2.modeling`, `sam2.modeling.sam`, `sam2.utils.misc`
    for pkg in ("sam2", "sam2.modeling", "sam2.modeling.sam", "sam2.utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []  # mark as package
            sys.modules[pkg] = m
    misc = types.ModuleType("sam2.utils.misc")

    def mask_to_box(*a, **k):  # only used by an out-of-scope point sampler (sam2_utils.py:16,176)
        raise NotImplementedError("stub: sam2.utils.misc.mask_to_box is outside the hot path")

    misc.mask_to_box = mask_to_box
    sys.modules["sam2.utils.misc"] = misc

    pe = _load("sam2.modeling.position_encoding", os.path.join(_MODELING, "position_encoding.py"))
    su = _load("sam2.modeling.sam2_utils", os.path.join(_MODELING, "sam2_utils.py"))
    tr = _load("sam2.modeling.sam.transformer", os.path.join(_MODELING, "sam", "transformer.py"))
    ma = _load("sam2.modeling.memory_attention", os.path.join(_MODELING, "memory_attention.py"))
    ls = _load("_ref_sam2_video_losses", os.path.join(REF_ROOT, "sam2_video", "model", "losses.py"))
    ns = types.SimpleNamespace(
        position_encoding=pe, sam2_utils=su, transformer=tr, memory_attention=ma, losses=ls,
        MemoryAttention=ma.MemoryAttention, MemoryAttentionLayer=ma.MemoryAttentionLayer,
        RoPEAttention=tr.RoPEAttention, Attention=tr.Attention,
        MultiStepMultiMasksAndIous=ls.MultiStepMultiMasksAndIous, BCECategoryLoss=ls.BCECategoryLoss,
    )
    _cache["ns"] = ns
    return ns


def load_sam2_base():
    """The reference's ``SAM2Base`` class (sam2_video/model/modeling/sam2_base.py), importable once the SAM heads'
    vendored modules are registered.  Used to pin oracle/bank_oracle.py: the unmodified
    ``_prepare_memory_conditioned_features`` is run with stub image encoder / memory encoder / memory attention."""
    load()
    if "sam2_base" not in _cache:
        _load("sam2.modeling.sam.prompt_encoder", os.path.join(_MODELING, "sam", "prompt_encoder.py"))
        _load("sam2.modeling.sam.mask_decoder", os.path.join(_MODELING, "sam", "mask_decoder.py"))
        _cache["sam2_base"] = _load("sam2.modeling.sam2_base", os.path.join(_MODELING, "sam2_base.py"))
    return _cache["sam2_base"].SAM2Base


def load_memory_encoder():
    """The reference's memory-encoder module (sam2_video/model/modeling/memory_encoder.py) and PositionEmbeddingSine."""
    ns = load()
    if "memory_encoder" not in _cache:
        _cache["memory_encoder"] = _load("sam2.modeling.memory_encoder", os.path.join(_MODELING, "memory_encoder.py"))
    return _cache["memory_encoder"], ns.position_encoding


def build_memory_encoder():
    """configs/sam2/sam2.1_hiera_t.yaml:62-85 through the reference's own constructors, sub-modules created in yaml order."""
    me, pe = load_memory_encoder()
    position_encoding = pe.PositionEmbeddingSine(num_pos_feats=64, normalize=True, scale=None, temperature=10000)
    mask_downsampler = me.MaskDownSampler(kernel_size=3, stride=2, padding=1)
    fuser = me.Fuser(layer=me.CXBlock(dim=256, kernel_size=7, padding=3, layer_scale_init_value=1e-6, use_dwconv=True), num_layers=2)
    return me.MemoryEncoder(out_dim=64, position_encoding=position_encoding, mask_downsampler=mask_downsampler, fuser=fuser)


def load_masks():
    """The reference's ``sam2_video/utils/masks.py`` (``merge_object_results_to_category``), loaded by path: the
    package ``__init__`` pulls in out-of-scope modules, the file itself needs only torch, numpy, cv2 and loguru."""
    if "masks" not in _cache:
        _cache["masks"] = _load("_ref_sam2_video_masks", os.path.join(REF_ROOT, "sam2_video", "utils", "masks.py"))
    return _cache["masks"]


def build_memory_attention(ns=None, dropout: float = 0.1, feat_sizes=(64, 64)):
    """Construct the stack with the kwargs of configs/sam2/sam2.1_hiera_t.yaml:29-60."""
    ns = ns or load()
    sa = ns.RoPEAttention(rope_theta=10000.0, feat_sizes=list(feat_sizes), embedding_dim=256,
                          num_heads=1, downsample_rate=1, dropout=dropout)
    ca = ns.RoPEAttention(rope_theta=10000.0, feat_sizes=list(feat_sizes), rope_k_repeat=True,
                          embedding_dim=256, num_heads=1, downsample_rate=1, dropout=dropout,
                          kv_in_dim=64)
    layer = ns.MemoryAttentionLayer(activation="relu", dim_feedforward=2048, dropout=dropout,
                                    pos_enc_at_attn=False, self_attention=sa, d_model=256,
                                    pos_enc_at_cross_attn_keys=True,
                                    pos_enc_at_cross_attn_queries=False, cross_attention=ca)
    return ns.MemoryAttention(d_model=256, pos_enc_at_input=True, layer=layer, num_layers=4)
