"""The oracle (oracle/*.py) against the fixtures generated from the reference itself (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import hd_oracle as hdo
from oracle import tokenpacker_oracle as tpo


ALL_SCALE_FACTORS = [2, 3, 4, 1, 6, 8, 12, 24]      # every divisor of 24 the reference constructor accepts (builder.py:51-52)


@pytest.mark.parametrize("s", ALL_SCALE_FACTORS)
def test_projector_matches_reference_fp32(golden_dir, s):
    g = np.load(os.path.join(golden_dir, f"projector_s{s}_h128.npz"))
    hidden, n = int(g["hidden"]), int(g["n"])
    params = tpo.make_params(hidden, seed=int(g["param_seed"]))
    x0, xm = tpo.make_inputs(n, seed=int(g["input_seed"]))
    # the seeded generators must reproduce the exact tensors the reference saw
    np.testing.assert_array_equal(x0[0, :3, :5], g["x0_probe"])
    np.testing.assert_array_equal(xm[-1, -3:, -5:], g["xm_probe"])
    np.testing.assert_array_equal(params["k_proj_1.0.weight"][:3, :5], g["w_probe"])
    out = tpo.tokenpacker_forward(params, x0, xm, s)            # float64 restatement
    assert out.shape == g["out"].shape == (n, (24 // s) ** 2, hidden)
    err = np.abs(out - g["out"]).max()
    assert err < 5e-6, err                                       # fp32 reference rounding only
    out32 = tpo.tokenpacker_forward(params, x0, xm, s, dtype=np.float32)
    assert np.abs(out32 - g["out"]).max() < 2e-5


@pytest.mark.parametrize("s", [2, 3, 4])
def test_projector_full_width_subsample(golden_dir, s):
    g = np.load(os.path.join(golden_dir, f"projector_s{s}_h4096_bf16in.npz"))
    params = {k: tpo.round_bf16(v) for k, v in tpo.make_params(4096, seed=int(g["param_seed"])).items()}
    x0, xm = tpo.make_inputs(1, seed=int(g["input_seed"]))
    out = tpo.tokenpacker_forward(params, tpo.round_bf16(x0), tpo.round_bf16(xm), s, dtype=np.float32)
    sub = out[0, ::int(g["row_stride"]), ::int(g["col_stride"])]
    assert np.abs(sub - g["out_sub"]).max() < 2e-5
    assert abs(float(np.sqrt((out.astype(np.float64) ** 2).mean())) - float(g["out_rms"])) < 1e-5


def test_bad_scale_factor():
    params = tpo.make_params(128)
    x0, xm = tpo.make_inputs(1)
    with pytest.raises(ValueError):
        tpo.tokenpacker_forward(params, x0, xm, 5)


def test_stencil_equivalence():
    """bilinear 24->g with align_corners=False is a fixed stencil per window (SURVEY.md key facts)."""
    x0, _ = tpo.make_inputs(1, seed=5)
    img = x0.astype(np.float64).reshape(1, 24, 24, -1)
    q2 = tpo.point_queries(x0.astype(np.float64), 2).reshape(1, 12, 12, -1)
    np.testing.assert_allclose(q2, img.reshape(1, 12, 2, 12, 2, -1).mean(axis=(2, 4)), atol=1e-12)
    q3 = tpo.point_queries(x0.astype(np.float64), 3).reshape(1, 8, 8, -1)
    np.testing.assert_allclose(q3, img[:, 1::3, 1::3], atol=1e-12)
    q4 = tpo.point_queries(x0.astype(np.float64), 4).reshape(1, 6, 6, -1)
    np.testing.assert_allclose(q4, img.reshape(1, 6, 4, 6, 4, -1)[:, :, 1:3, :, 1:3].mean(axis=(2, 4)), atol=1e-12)


def test_window_locality_property():
    """Perturbing a fine token outside query m's window must not change output row m."""
    s, hidden = 4, 64
    params = tpo.make_params(hidden, seed=3)
    x0, xm = tpo.make_inputs(1, seed=4)
    base = tpo.tokenpacker_forward(params, x0, xm, s)
    xm2 = xm.copy()
    tok = 5 * 24 + 9            # fine token (row 5, col 9) -> window (hb=1, wb=2) -> query 1*6+2 = 8
    xm2[0, tok] += 1.0
    pert = tpo.tokenpacker_forward(params, x0, xm2, s)
    changed = np.abs(pert - base).max(axis=-1)[0] > 1e-9
    assert changed[8] and changed.sum() == 1


def test_round_bf16_matches_torch():
    import torch
    a = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 3
    ref = torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    np.testing.assert_array_equal(tpo.round_bf16(a), ref)


def test_flop_model():
    assert abs(tpo.flops_per_crop(2) / 1e9 - 21.4436) < 1e-3
    assert abs(tpo.flops_per_crop(3) / 1e9 - 17.5849) < 1e-3
    assert abs(tpo.flops_per_crop(4) / 1e9 - 16.2343) < 1e-3
    assert tpo.bytes_per_crop(2) == 7077888


# ------------------------------------------------------------------ HD front end

def test_hd_grid_table(golden_dir):
    table = np.load(os.path.join(golden_dir, "hd_grid.npz"))["table"]
    bad = [(h, w, p) for h, w, p, hb, wb in table.tolist() if hdo.hd_grid(h, w, p) != (hb, wb)]
    assert not bad, bad[:5]


def test_hd_grid_rejects_unknown_patch_num():
    with pytest.raises(NotImplementedError):
        hdo.hd_grid(336, 336, 12)


def test_hd_tile(golden_dir):
    g = np.load(os.path.join(golden_dir, "hd_tile.npz"))
    for ci in range(int(g["n_cases"])):
        h, w, patch_num, hb, wb, seed = (int(v) for v in g[f"case{ci}_meta"])
        img = np.random.default_rng(seed).standard_normal((3, h, w)).astype(np.float32)
        crops, ohb, owb = hdo.hd_tile(img[None], patch_num)
        assert (ohb, owb) == (hb, wb)
        assert crops.shape == (hdo.n_crops(hb, wb), 3, 336, 336)
        np.testing.assert_allclose(crops[:, :, ::37, ::41], g[f"case{ci}_probe"], atol=2e-6)
        np.testing.assert_allclose(crops.astype(np.float64).sum(axis=(1, 2, 3)), g[f"case{ci}_sum"], atol=5e-2)
        np.testing.assert_allclose(np.abs(crops.astype(np.float64)).sum(axis=(1, 2, 3)), g[f"case{ci}_abs"], rtol=1e-6)


def test_hd_assemble(golden_dir):
    g = np.load(os.path.join(golden_dir, "hd_assemble.npz"))
    grids = g["grids"].tolist()
    packed, cu = hdo.hd_assemble(g["feats"], [a for a, _ in grids], [b for _, b in grids], g["sep_row"], g["ret_row"])
    np.testing.assert_array_equal(cu, g["cu"])
    np.testing.assert_array_equal(packed, g["packed"])
    assert hdo.hd_seq_len(3, 3, 144) == 1450 and hdo.hd_seq_len(1, 1, 144) == 145 and hdo.hd_seq_len(5, 5, 36) == 962


@pytest.mark.parametrize("s", ALL_SCALE_FACTORS)
def test_torch_port_matches_reference(golden_dir, s):
    """The PyTorch-CPU port that bench.py times as the CPU baseline reproduces the reference's own output."""
    import torch
    from oracle import torch_port
    g = np.load(os.path.join(golden_dir, f"projector_s{s}_h128.npz"))
    params = {k: torch.from_numpy(v) for k, v in tpo.make_params(int(g["hidden"]), seed=int(g["param_seed"])).items()}
    x0, xm = tpo.make_inputs(int(g["n"]), seed=int(g["input_seed"]))
    with torch.no_grad():
        out = torch_port.forward(params, torch.from_numpy(x0), torch.from_numpy(xm), s).numpy()
    assert np.abs(out - g["out"]).max() < 1e-6
