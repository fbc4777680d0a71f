"""K5: the mask-pyramid kernels (csrc/mask_plan_kernels.cu, smb_view_level_masks / smb_view_level_plan)
against the reference's own torch ops (tests/fake_engine.py restates model/model.py:204-254 and
content_and_style_losses.py:161,172-185 with F.conv2d / F.interpolate on the CPU):

  * level masks / depth-interpolation weights at the rgb resolution          bit-exact
  * nearest-resampled row masks of every VGG layer, their pixel counts        bit-exact (index math of UpSample.h)
  * the angle pass / fail split                                               exact up to pixels whose bilinear angle is
                                                                              within 1 ulp of the threshold
  * bilinear angle hook                                                       <= 2 ulp (FMA contraction on the CPU side)
Sizes: the reference's ScanNet / Matterport pyramids (up-sampling 256 -> 784, odd widths 341 / 811 / 1045), a
down-sampling level, identity and exact-2x levels (the two shortcuts of nearest_idx)."""
import pytest
import torch

import fake_engine

pytestmark = pytest.mark.gpu


def _layer_sizes(H, W):
    from stylemesh_b200.model.losses.content_and_style_losses import layer_hw
    return [layer_hw(c, H, W) for c in (0, 2, 4, 8, 12, 9)]


@pytest.mark.parametrize("rgb,levels", [((256, 341), [(256, 341), (432, 576), (608, 811), (784, 1045)]),
                                        ((256, 320), [(256, 320), (432, 540), (608, 760), (784, 980)]),
                                        ((96, 131), [(48, 64), (96, 131), (192, 262), (133, 77)])])
def test_level_masks_and_plans_match_the_reference_ops(rgb, levels):
    from stylemesh_b200 import engine as eng, synthetic as syn
    v = syn.make_view(77, rgb, levels)
    Hr, Wr = rgb
    L = len(levels)
    mask, r, o, w = v.mask[0], v.rounded_depth_level[0, 0], v.other_depth_level[0, 0], v.interp_weight[0, 0]
    want_m, want_w = fake_engine.view_level_masks(mask, r, o, w, L)
    got_m, got_w = eng.view_level_masks(mask.cuda(), r.cuda(), o.cuda(), w.cuda(), L)
    assert torch.equal(got_m.cpu(), want_m) and torch.equal(got_w.cpu(), want_w)
    assert float(want_m.sum()) > 0                                  # the erosion leaves something to compare

    thr = 30.0
    guid, deg = v.angle_guidance[0, 0].contiguous(), v.angle_degrees[0, 0].contiguous()
    for i, (H, W) in enumerate(levels):
        lh = _layer_sizes(H, W)
        cw = torch.zeros(1 + 3 * len(lh), dtype=torch.int32)
        want = fake_engine.view_level_plan(want_m[i], want_w[i], guid, deg, thr, (H, W), lh, cw)
        cg = torch.zeros(1 + 3 * len(lh), dtype=torch.int32, device="cuda")
        got = eng.view_level_plan(got_m[i], got_w[i], guid.cuda(), deg.cuda(), thr, (H, W), lh, cg)
        assert torch.equal(got["hook1"].cpu(), want["hook1"]), (i, "depth-interpolation hook")
        assert torch.allclose(got["hook0"].cpu(), want["hook0"], rtol=3e-7, atol=1e-7), (i, "angle hook")
        cg = cg.cpu()
        assert int(cg[0]) == int(cw[0]), (i, "alive pixels")
        for k, (a, b) in enumerate(zip(got["layers"], want["layers"])):
            assert torch.equal(a["mask"].cpu(), b["mask"]), (i, k)
            assert int(cg[1 + 3 * k]) == int(cw[1 + 3 * k]) == int(b["mask"].sum())
            for key in ("mask_pass", "mask_fail"):
                assert int((a[key].cpu() != b[key]).sum()) <= 2, (i, k, key)          # pixels exactly at the threshold
            assert abs(int(cg[2 + 3 * k]) - int(cw[2 + 3 * k])) <= 2 and abs(int(cg[3 + 3 * k]) - int(cw[3 + 3 * k])) <= 2
            assert torch.equal(a["mask_pass"] + a["mask_fail"], a["mask"])             # the split partitions the mask
            assert int(cg[2 + 3 * k]) + int(cg[3 + 3 * k]) == int(cg[1 + 3 * k])


def test_plain_mask_without_depth_levels_or_split():
    """use_depth_scaling=False: the last level takes nearest(mask) > 0 (model.py:253-254); no hooks, no split."""
    from stylemesh_b200 import engine as eng, synthetic as syn
    v = syn.make_view(78, (120, 160), [(120, 160)])
    m = v.mask[0].float().contiguous()
    for (H, W) in [(120, 160), (240, 320), (171, 211)]:
        lh = _layer_sizes(H, W)
        cw = torch.zeros(1 + 3 * len(lh), dtype=torch.int32)
        want = fake_engine.view_level_plan(m, None, None, None, 0.0, (H, W), lh, cw)
        cg = torch.zeros(1 + 3 * len(lh), dtype=torch.int32, device="cuda")
        got = eng.view_level_plan(m.cuda(), None, None, None, 0.0, (H, W), lh, cg)
        assert got["hook0"] is None and got["hook1"] is None
        assert torch.equal(cg.cpu(), cw)
        for a, b in zip(got["layers"], want["layers"]):
            assert "mask_pass" not in a and torch.equal(a["mask"].cpu(), b["mask"])
