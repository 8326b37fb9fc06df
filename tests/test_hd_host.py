"""Host-side HD front end (no CUDA): grid selector and slice-assembly plan vs the reference-generated fixtures."""
import os

import numpy as np
import pytest

from oracle import hd_oracle as hdo


def test_grid_selector_matches_reference_table(golden_dir):
    from tokenpacker_b200 import hd_grid, Image_Patch
    table = np.load(os.path.join(golden_dir, "hd_grid.npz"))["table"]
    bad = [(h, w, p) for h, w, p, hb, wb in table.tolist() if hd_grid(h, w, p) != (hb, wb)]
    assert not bad, bad[:5]
    assert Image_Patch(image_size=336, patch_num=9).calculate(1088, 1088) == (3, 3)
    with pytest.raises(NotImplementedError):
        Image_Patch(patch_num=12)
    with pytest.raises(NotImplementedError):
        hd_grid(100, 100, 7)


def test_fit_matches_oracle():
    from tokenpacker_b200.hd import hd_fit, hd_grid
    rng = np.random.default_rng(3)
    for _ in range(300):
        h, w = (int(v) for v in rng.integers(20, 3000, size=2))
        hb, wb = hd_grid(h, w, 16)
        main, thumb = hd_fit(h, w, hb, wb)
        assert main == hdo._fit(h, w, hb, wb) and thumb == hdo._fit(h, w, 1, 1)


def test_plan_matches_reference_assembly(golden_dir):
    from tokenpacker_b200 import hd_plan, hd_seq_len
    g = np.load(os.path.join(golden_dir, "hd_assemble.npz"))
    grids = g["grids"].tolist()
    m = g["feats"].shape[1]
    plan = hd_plan([a for a, _ in grids], [b for _, b in grids], m)
    np.testing.assert_array_equal(plan.cu_seqlens.numpy(), g["cu"])
    assert plan.n_crops == g["feats"].shape[0]
    # rebuild the packed sequence from the plan and compare with the reference's torch.cat result
    packed = np.full_like(g["packed"], np.nan)
    for c, r0 in enumerate(plan.seg_row_offset.tolist()):
        packed[r0:r0 + m] = g["feats"][c]
    packed[plan.sep_rows.numpy()] = g["sep_row"]
    packed[plan.ret_rows.numpy()] = g["ret_row"]
    np.testing.assert_array_equal(packed, g["packed"])
    for (a, b), l0, l1 in zip(grids, g["cu"][:-1], g["cu"][1:]):
        assert l1 - l0 == hd_seq_len(a, b, m) == hdo.hd_seq_len(a, b, m)
    assert hd_seq_len(3, 3, 144) == 1450
