"""Pins the conclusions of tools/vis_numerics_study.py (CPU emulation of the tensor-core engine's number formats) that
the GPU tolerances rest on: (1) the scaled fp16 hi/lo 3-term split of round 2 is as accurate as the fp32 reference path,
forward and backward, and flips no ReLU unit; (2) the bf16 split of round 1 was 10x less accurate forward and its
gradient error was ReLU units taking the other branch -- with the fp64 masks forced the same backward agrees to 1e-5;
(3) the single-pass fast mode is a 1e-4-class forward with a percent-level direction gradient."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_number_formats_of_the_visibility_engine():
    import vis_numerics_study as S
    rows, units = S.study(n=6000)
    fp32, tc = rows["fp32"], rows["fp16 hi/lo x3, scaled (engine tc)"]
    bf, tc1 = rows["bf16 hi/lo x3 (round 1)"], rows["fp16 single pass (engine tc1)"]
    # forward: the scaled fp16 split is fp32-class, the bf16 split 10x worse, the fast mode 1e-4-class
    assert fp32["vis_max_abs"] < 5e-7 and tc["vis_max_abs"] < 5e-7
    assert 3 * tc["vis_max_abs"] < bf["vis_max_abs"] < 1e-5
    assert 1e-5 < tc1["vis_max_abs"] < 1e-3
    # backward with the fp64 ReLU masks forced = operand precision alone
    assert fp32["grad_rel_l2_fp64_masks"] < 5e-6 and tc["grad_rel_l2_fp64_masks"] < 5e-6
    assert bf["grad_rel_l2_fp64_masks"] < 3e-5 and tc1["grad_rel_l2_fp64_masks"] < 5e-3
    # free-running backward: identical to the forced one when no unit flipped; ONE flipped unit among millions moves the
    # relative L2 of the whole gradient to ~1e-3 (it happens to the exact fp32 path too: this sample has one)
    for r in rows.values():
        if r["flips"] == 0:
            assert abs(r["grad_rel_l2"] - r["grad_rel_l2_fp64_masks"]) < 1e-7
        else:
            assert r["grad_rel_l2"] > 5 * r["grad_rel_l2_fp64_masks"]
    assert tc["flips"] <= fp32["flips"] + 1 and bf["flips"] > tc["flips"] and tc1["flips"] > 20 * max(bf["flips"], 1)
