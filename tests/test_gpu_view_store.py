"""GPU parity of the view-preparation kernels and the resident view store against the 13-tuples the
REAL reference dataset class produced for a synthetic ScanNet-layout scene (tests/golden/view_prep.npz), plus
size-independent properties at the full benchmark size and the CLI running `--dataset scannet` end to end."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import view_scene_util as vsu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(vsu.GOLD)


def test_store_matches_reference_tuples(gold, tmp_path):
    from stylemesh_b200.data.scannet_scene import ScanNetScene, load_scene_into_store
    root = vsu.write_scene(gold, tmp_path)
    sc = ScanNetScene(f"{root}/train/images/{vsu.SCENE}", pyramid_levels=3, min_pyramid_height=32)
    store = load_scene_into_store(sc, "cuda", 30, min_pyramid_depth=1.0)
    assert len(store) == 3 and store.bytes_resident > 0
    for i in range(3):
        v = store[i]
        assert all(t.is_cuda for j, t in enumerate(v) if j not in (8, 9)) and all(u.is_cuda for u in v[9])
        vsu.check_view_against_golden(v, gold, i)


def test_matterport_store_matches_reference_tuples(tmp_path):
    from stylemesh_b200.data.matterport_scene import MatterportRegion
    from stylemesh_b200.data.scene_base import load_scene_into_store
    mg = np.load(vsu.MP_GOLD)
    root = vsu.write_matterport(mg, tmp_path)
    reg = MatterportRegion(f"{root}/v1/scans/{vsu.MP_HOUSE}", region_index=0, pyramid_levels=3, min_pyramid_height=32)
    store = load_scene_into_store(reg, "cuda", 30, min_pyramid_depth=1.0)
    for i in range(3):
        vsu.check_view_against_golden(store[i], mg, i)


def test_rendered_depth_store_matches_reference_tuples(tmp_path):
    from stylemesh_b200.data.scannet_scene import ScanNetScene, load_scene_into_store
    rg = np.load(vsu.RD_GOLD)
    root = vsu.write_rendered_depth_scene(rg, tmp_path)
    sc = ScanNetScene(f"{root}/train/images/{vsu.RD_SCENE}", pyramid_levels=2, min_pyramid_height=32)
    store = load_scene_into_store(sc, "cuda", 30, min_pyramid_depth=1.0)
    for i in range(2):
        vsu.check_view_against_golden(store[i], rg, i)


def test_kernels_equal_the_oracle_on_ragged_sizes():
    """Each kernel against the pinned numpy oracle on sizes that are not multiples of anything; integer outputs,
    gathers and the explicitly rounded float arithmetic must be bit-identical."""
    from oracle import view_prep_oracle as vo
    from stylemesh_b200 import engine as eng
    from stylemesh_b200.data import resample as rs
    rng = np.random.default_rng(3)
    dev = "cuda"
    # depth: uint16 sensor map -> working size, depth levels
    depth_mm = rng.integers(0, 6000, size=(97, 131), dtype=np.uint16)
    depth_mm[rng.random((97, 131)) < 0.1] = 0
    for (hd, wd) in [(61, 83), (97, 131), (200, 301)]:
        yo, ya = rs.cv2_linear_table(97, hd)
        xo, xa = rs.cv2_linear_table(131, wd)
        tabs = tuple(torch.from_numpy(t).to(dev) for t in (yo, ya, xo, xa))
        got = eng.view_resize_linear(torch.from_numpy(depth_mm).to(dev), (hd, wd), None if (hd, wd) == (97, 131) else tabs,
                                     1000.0).cpu().numpy()
        want = vo.resize_linear_cv2(depth_mm / 1000.0, (wd, hd))
        assert np.array_equal(got, want)                                  # float64, same operations in the same order
        levels = [256.0, 432.0, 608.0, 784.0]
        c, d32, r, o, w = [t.cpu().numpy() for t in eng.view_depth_levels(torch.from_numpy(want).to(dev), levels, 0.25)]
        wc, wr, wo, ww = vo.depth_levels(want, levels, 0.25)
        assert np.array_equal(r, wr) and np.array_equal(o, wo)
        assert np.array_equal(c, wc) and np.array_equal(w, ww) and np.array_equal(d32, want.astype(np.float32))
    # float32 (rendered) depth keeps numpy's float32 arithmetic
    d32 = (rng.random((40, 50)) * 4).astype(np.float32)
    c, _, r, o, w = [t.cpu().numpy() for t in eng.view_depth_levels(torch.from_numpy(d32.astype(np.float64)).to(dev),
                                                                     [32.0, 48.0, 64.0], 0.5, depth_is_f32=True)]
    wc, wr, wo, ww = vo.depth_levels(d32, [32.0, 48.0, 64.0], 0.5)
    assert np.array_equal(r, wr) and np.array_equal(o, wo) and np.array_equal(c, wc) and np.array_equal(w, ww)
    # uv -> grid + mask (with and without the depth factor)
    uv = rng.random((77, 53, 3)).astype(np.float32)
    uv[rng.random((77, 53)) < 0.2] = 0
    dep = rng.random((77, 53)) - 0.3
    g, m = eng.view_uv_to_grid(torch.from_numpy(uv).to(dev), want_mask=True, depth_at_uv=torch.from_numpy(dep).to(dev))
    assert np.array_equal(g.cpu().numpy(), vo.uv_to_grid(uv))
    assert np.array_equal(m.cpu().numpy(), ((uv[:, :, 0] != 0) | (uv[:, :, 1] != 0)) & (dep > 0))
    g2, m2 = eng.view_uv_to_grid(torch.from_numpy(uv).to(dev), want_mask=True)
    assert np.array_equal(m2.cpu().numpy(), vo.uv_valid_mask(uv))
    # gathers, rgb, angle, erosion
    a = rng.random((48, 64)).astype(np.float32)
    yt, xt = [torch.from_numpy(rs.cv2_nearest_table(s, d)).to(dev) for s, d in ((48, 37), (64, 91))]
    assert np.array_equal(eng.view_gather2d(torch.from_numpy(a).to(dev), yt, xt).cpu().numpy(),
                          vo.resize_nearest_cv2(a, (91, 37)))
    mk = rng.random((48, 64)) > 0.5
    yt, xt = [torch.from_numpy(rs.pil_nearest_table(s, d)).to(dev) for s, d in ((48, 37), (64, 91))]
    assert np.array_equal(eng.view_gather2d(torch.from_numpy(mk.astype(np.uint8)).to(dev), yt, xt).cpu().numpy() > 0,
                          vo.resize_nearest_pil(mk, (91, 37)))
    rgb = rng.integers(0, 256, size=(33, 45, 3), dtype=np.uint8)
    assert np.array_equal(eng.view_rgb_pre(torch.from_numpy(rgb).to(dev)).cpu().numpy(), vo.rgb_pre(rgb))
    cosang = (rng.random((33, 45)) * 2 - 1).astype(np.float32)
    assert np.allclose(eng.view_angle_degrees(torch.from_numpy(cosang).to(dev)).cpu().numpy(), vo.angle_degrees(cosang),
                       rtol=0, atol=3e-5)
    x = (rng.random((1, 1, 41, 59)) > 0.15).astype(np.float32)
    xt_ = torch.from_numpy(x)
    e = torch.clamp(F.conv2d(xt_, torch.ones(1, 1, 3, 3), padding=1) / 9, 0, 1)
    assert torch.equal(eng.view_erode3x3(xt_.cuda()).cpu(), xt_ * (e == 1))          # model.py:204-208


def test_full_size_properties():
    """640x480 working size, 784-row UV level: idempotence of same-size resampling, levels inside the table, weights in
    (0.5, 1], masks that only shrink under the depth factor, grid range."""
    from stylemesh_b200.data import RawView, ViewStore
    rng = np.random.default_rng(11)
    H, W = 480, 640
    sizes = [(256, 341), (432, 576), (608, 811), (784, 1045)]
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depth = (500 + 3500 * ((xx / W + yy / H) % 1.0)).astype(np.uint16)
    depth[:, :7] = 0
    uvs = []
    for (h, w) in sizes:
        uv = rng.random((h, w, 3)).astype(np.float32) * 0.98 + 0.01
        uv[: h // 9] = 0
        uvs.append(uv)
    raw = RawView(rgb=rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8), uv_pyramid=uvs,
                  angle=rng.random((H, W, 3)).astype(np.float32), depth=depth, depth_divisor=1000.0)
    store = ViewStore("cuda", [256, 432, 608, 784], 0.25, (W, H))
    v = store[store.add(raw)]
    rgb, _, _, d, level, r, o, w, idx, grids, mask, ang, ang_deg = v
    assert torch.equal(d[0, 0].cpu(), torch.from_numpy((depth / 1000.0).astype(np.float32)))     # same size: conversion only
    assert int(r.min()) >= 0 and int(r.max()) <= 3 and int(o.min()) >= 0 and int(o.max()) <= 3
    assert int((r - o).abs().max()) <= 1
    assert float(w.min()) > 0.5 - 1e-6 and float(w.max()) <= 1.0
    assert bool(((level - o.float()).abs() <= 1.0 + 1e-6).all())
    for g, (h, wd) in zip(grids, sizes):
        assert g.shape == (1, h, wd, 2) and float(g.min()) >= -1.0 and float(g.max()) <= 1.0
        assert bool((g[0, : h // 9] == -1).all())                          # invalid pixels sit exactly at (-1, -1)
    store2 = ViewStore("cuda", [256, 432, 608, 784], 0.25, (W, H), mask_uses_depth=False)
    mask2 = store2[store2.add(raw)][10]
    assert bool((mask <= mask2).all()) and int(mask2.sum()) > int(mask.sum()) > 0
    assert torch.equal(ang[0, 0].cpu(), torch.from_numpy(raw.angle[:, :, 0]))                     # same size: identity gather
    assert float(ang_deg.min()) >= 0.0 and float(ang_deg.max()) <= 90.0 + 1e-3


def test_cli_runs_a_scannet_scene_from_disk(gold, tmp_path):
    """`python -m model.optimize --dataset scannet` (model/optimize.py:44-63) on the synthetic scene: reader -> view
    store -> two epochs of the fused step; the loss stays finite and the texture moves."""
    from stylemesh_b200.model.optimize import build_parser, main
    root = vsu.write_scene(gold, tmp_path / "data")
    args = build_parser().parse_args([
        "--style_image_path", "synthetic:96:80", "--vgg_gatys_model_path", "synthetic:0", "--root_path", root,
        "--dataset", "scannet", "--scene", vsu.SCENE, "--resize_size", "30", "--pyramid_levels", "3",
        "--min_pyramid_height", "32", "--min_pyramid_depth", "1.0", "--min_images", "1", "--max_images", "10",
        "--texture_size", "64,64", "--hierarchical", "--hierarchical_layers", "2", "--random_texture_init",
        "--loss_weight", "content=70", "--loss_weight", "style=1e-4", "--loss_weight", "tex_reg=5e3",
        "--train_split", "0.7", "--sampler_mode", "repeat", "--index_repeat", "2", "--max_epochs", "2",
        "--default_root_dir", str(tmp_path / "logs")])
    model = main(args)
    tex = torch.cat([l.data.detach().reshape(-1) for l in model.texture.layers])
    assert bool(torch.isfinite(tex).all()) and float(tex.abs().max()) > 0
