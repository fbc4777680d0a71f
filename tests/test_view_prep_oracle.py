"""Pins oracle/view_prep_oracle.py against tests/golden/view_prep.npz (the real reference dataset class run on a
synthetic ScanNet-layout scene by tests/golden/make_view_golden.py).  Integer / index / mask outputs are compared
bit for bit; float outputs exactly where the reference's arithmetic is restated operation by operation, with a stated
tolerance where a third-party transcendental is involved (torch.acos vs numpy.arccos)."""
import os

import numpy as np
import pytest

from oracle import view_prep_oracle as vo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "view_prep.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _views(gold):
    n, resize, nlev = [int(x) for x in gold["meta"]]
    return n, resize, nlev


def _prep(gold, i):
    from PIL import Image
    n, resize, nlev = _views(gold)
    rgb = gold[f"rgb_{i}"]
    size_wh = vo.resolve_resize(resize, (rgb.shape[1], rgb.shape[0]))
    rgb_r = np.asarray(Image.fromarray(rgb).resize(size_wh))                 # abstract_dataset.py:299 (PIL default filter)
    levels = gold["levels"]
    uv = [gold[f"uv{int(h)}_{i}"] for h in levels]
    depth = np.asarray(gold[f"depth_mm_{i}"]) / 1000.0                        # scannet_dataset.py:301
    return vo.preprocess_view(rgb_r, uv, gold[f"angle_{i}"], depth, levels, 1.0, size_wh), size_wh


def test_level_selection_matches_the_reference_folder_filter(gold):
    # uv_24 is below min_pyramid_height = 32; pyramid_levels = 3 keeps 32, 48, 64 (scannet_dataset.py:225-236)
    assert gold["levels"].tolist() == [32.0, 48.0, 64.0]
    assert gold["all_levels"].tolist() == [24.0, 32.0, 48.0, 64.0]


@pytest.mark.parametrize("i", [0, 1, 2])
def test_integer_and_mask_outputs_bit_exact(gold, i):
    out, _ = _prep(gold, i)
    assert np.array_equal(out["mask"], gold[f"ref_mask_{i}"])
    assert np.array_equal(out["rounded_depth_level"], gold[f"ref_rounded_depth_level_{i}"])
    assert np.array_equal(out["other_depth_level"], gold[f"ref_other_depth_level_{i}"])
    assert out["rounded_depth_level"].dtype == np.int64


@pytest.mark.parametrize("i", [0, 1, 2])
def test_float_outputs(gold, i):
    out, _ = _prep(gold, i)
    for l in range(3):
        assert np.array_equal(out["uv"][l], gold[f"ref_uv{l}_{i}"])                      # fl(fl(2u) - 1): exact
    assert np.array_equal(out["rgb"], gold[f"ref_rgb_{i}"])                                # ToTensor + pre(): exact
    assert np.array_equal(out["angle_guidance"], gold[f"ref_angle_guidance_{i}"])          # nearest gather: exact
    # cv2's 2x-downscale fast path averages (a+b+c+d)*0.25 instead of two lerps: <= 1 ulp of the double result
    assert np.allclose(out["depth"], gold[f"ref_depth_{i}"], rtol=1e-6, atol=0)
    assert np.allclose(out["depth_level"], gold[f"ref_depth_level_{i}"], rtol=0, atol=2e-6)
    assert np.allclose(out["interp_weight"], gold[f"ref_interp_weight_{i}"], rtol=0, atol=2e-6)
    # torch.acos (sleef) vs numpy.arccos: a few float32 ulps at 90 degrees
    assert np.allclose(out["angle_degrees"], gold[f"ref_angle_degrees_{i}"], rtol=0, atol=2e-5)


def test_resampling_tables_against_the_libraries():
    """The three resampling conventions, pinned directly against cv2 / PIL on ramps (ragged up- and down-scales)."""
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    rng = np.random.default_rng(0)
    for (hs, ws), (hd, wd) in [((48, 64), (30, 40)), ((60, 80), (32, 42)), ((33, 47), (64, 85)), ((64, 85), (30, 40)),
                               ((60, 80), (48, 64))]:
        img = rng.random((hs, ws))
        want = cv2.resize(img, (wd, hd), interpolation=cv2.INTER_LINEAR)
        assert np.allclose(vo.resize_linear_cv2(img, (wd, hd)), want, rtol=1e-12, atol=1e-15)
        img32 = img.astype(np.float32)
        want32 = cv2.resize(img32, (wd, hd), interpolation=cv2.INTER_LINEAR)
        assert np.allclose(vo.resize_linear_cv2(img32, (wd, hd)), want32, rtol=2e-6, atol=1e-7)
        assert np.array_equal(vo.resize_nearest_cv2(img32, (wd, hd)),
                              cv2.resize(img32, (wd, hd), interpolation=cv2.INTER_NEAREST))
        m = rng.random((hs, ws)) > 0.5
        want_m = np.asarray(Image.fromarray(m).resize((wd, hd), Image.NEAREST))
        assert np.array_equal(vo.resize_nearest_pil(m, (wd, hd)), want_m)


def test_intrinsics_rescale(gold):
    k = np.identity(4, dtype=np.float32)
    k[0, 0], k[1, 1], k[0, 2], k[1, 2] = 70.5, 71.25, 39.5, 29.5
    got = vo.modify_intrinsics(k, (80, 60), (40, 30))
    assert np.array_equal(got, gold["ref_intrinsics_0"])


@pytest.mark.parametrize("i", [0, 1, 2])
def test_matterport_variant(i):
    """Depth in 1/4000 m (matterport_dataset.py:288) and a validity mask that ignores the depth (:304-307), against the
    reference's Matterport_Single_House_Dataset on a synthetic house region."""
    from PIL import Image
    g = np.load(os.path.join(os.path.dirname(GOLD), "view_prep_matterport.npz"))
    rgb = g[f"mp_rgb_{i}"]
    size_wh = vo.resolve_resize(int(g["meta"][1]), (rgb.shape[1], rgb.shape[0]))
    rgb_r = np.asarray(Image.fromarray(rgb).resize(size_wh))
    uv = [g[f"mp_uv{int(h)}_{i}"] for h in g["levels"]]
    out = vo.preprocess_view(rgb_r, uv, g[f"mp_angle_{i}"], np.asarray(g[f"mp_depth_q_{i}"]) / 4000.0, g["levels"], 1.0,
                             size_wh, mask_uses_depth=False)
    assert np.array_equal(out["mask"], g[f"ref_mask_{i}"])
    assert np.array_equal(out["rounded_depth_level"], g[f"ref_rounded_depth_level_{i}"])
    assert np.array_equal(out["other_depth_level"], g[f"ref_other_depth_level_{i}"])
    assert np.array_equal(out["rgb"], g[f"ref_rgb_{i}"]) and np.array_equal(out["uv"][2], g[f"ref_uv2_{i}"])
    assert np.allclose(out["depth"], g[f"ref_depth_{i}"], rtol=1e-6, atol=0)
    assert np.allclose(out["interp_weight"], g[f"ref_interp_weight_{i}"], rtol=0, atol=2e-6)
