"""Host logic of the memory-bank assembly (no GPU): frame / pointer selection of sam2_video_training_b200.memory_bank
against the oracle restatement of sam2_base.py:551-647 (which tests/test_oracle_golden.py pins to the reference)."""
import pytest
import torch

from oracle import bank_oracle as bo
from oracle import detgen
from sam2_video_training_b200 import memory_bank as mb


@pytest.mark.parametrize("tag", [t for t, _ in detgen.bank_scenarios()])
def test_selection_matches_oracle(tag):
    kw = dict(detgen.bank_scenarios())[tag]
    od, _, _, _ = detgen.bank_inputs(kw["cond"], kw["non_cond"])
    args = dict(max_cond_frames_in_attn=kw.get("max_cond", -1), memory_temporal_stride_for_eval=kw.get("stride", 1))
    frames, pointers = mb.select_bank_entries(mb.BankConfig(**args), kw["frame_idx"], od, kw["num_frames"], kw["training"],
                                              kw.get("reverse", False))
    ocfg = bo.BankConfig(**args)
    tp, sel, unsel = bo.select_memory_frames(ocfg, kw["frame_idx"], od, kw["training"], kw.get("reverse", False))
    want_frames = [(t, o) for t, o in tp if o is not None]
    assert [(t, id(o)) for t, o in frames] == [(t, id(o)) for t, o in want_frames]
    want_ptrs = bo.select_object_pointers(ocfg, kw["frame_idx"], od, kw["num_frames"], kw["training"], sel, unsel, kw.get("reverse", False))
    assert [(d, id(p)) for d, p in pointers] == [(d, id(p)) for d, p in want_ptrs]


def test_select_closest_cond_frames_and_sine_pe():
    outs = {t: {"k": t} for t in (0, 3, 7, 14, 20)}
    for frame_idx in (1, 7, 10, 25):
        for k in (-1, 2, 3, 5):
            a, b = mb.select_closest_cond_frames(frame_idx, outs, k)
            c, d = bo.select_closest_cond_frames(frame_idx, outs, k)
            assert list(a) == list(c) and list(b) == list(d)
    x = torch.tensor([0.0, 0.2, -0.5, 1.0])
    assert torch.equal(mb.get_1d_sine_pe(x, 256), bo.get_1d_sine_pe(x, 256))


def test_cpu_tensors_are_rejected():
    od, tpos, pw, pb = detgen.bank_inputs([0], [1, 2])
    from sam2_video_training_b200 import _lib
    with pytest.raises(_lib.Sam2B200Error):
        mb.assemble_memory(mb.BankConfig(), 3, od, 8, tpos, None, training=True)
