"""The CUDA rasteriser (csrc/raster_kernels.cu, smb_raster_view) against the CPU oracle of the reference's OpenGL renderer
(oracle/raster_oracle.py), and the whole headless chain it enables: mesh -> UV / angle / depth maps in the reference's
ScanNet layout -> stylemesh_b200.data.ScanNetScene -> a training step -> mip-mapped preview."""
import os

import numpy as np
import pytest
import torch

import raster_scene_util as rs
from oracle import raster_oracle as ro

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wh,flip", [((341, 256), False), ((160, 120), True), ((1045, 784), False)])
def test_rasteriser_matches_the_oracle(wh, flip):
    from stylemesh_b200 import raster
    verts, faces, cuv, cn = rs.room_mesh()
    r = raster.MeshRasterizer(raster.Mesh(verts, faces, cuv, cn))
    for pose in rs.room_poses(3):
        want = ro.render(verts, faces, cuv, cn, pose, rs.INTRINSICS, rs.INTRINSICS_SIZE, wh, flip=flip)
        got = [t.cpu().numpy() for t in r.render(pose, rs.INTRINSICS, rs.INTRINSICS_SIZE, wh, flip=flip)]
        for g, w_ in zip(got, want):
            assert g.shape == w_.shape == (wh[1], wh[0], 3) and g.dtype == np.float32
        cov_g, cov_w = got[2][..., 0] > 0, want[2][..., 0] > 0
        assert np.mean(cov_g != cov_w) < 2e-3                               # coverage: ties on triangle edges only
        # same surface point wherever both see the same triangle (depth agrees): attributes to float32 accuracy
        same = cov_g & cov_w & (np.abs(got[2][..., 0] - want[2][..., 0]) < 1e-3)
        assert same.mean() > 0.97
        assert np.abs(got[2][..., 0] - want[2][..., 0])[same].max() < 1e-4
        assert np.abs(got[0][..., :2] - want[0][..., :2])[same].max() < 2e-4   # u, v
        assert np.abs(got[1][..., 0] - want[1][..., 0])[same].max() < 2e-4     # cos(view angle)
        assert np.abs(got[0][..., 2] - want[0][..., 2])[same].max() < 2e-2     # LOD (log2 of a ratio of small differences)
        assert np.array_equal(got[1][..., 0], got[1][..., 2]) and np.array_equal(got[2][..., 0], got[2][..., 1])
        assert float(np.abs(got[0][~cov_g]).max(initial=0)) == 0            # clear colour where there is no geometry


def test_mesh_to_training_step_without_opengl(tmp_path):
    """render_scene writes the reference's ScanNet layout (uv/, uv_<h>/, rendered depth behind an empty depth/ folder);
    the scene reader, the view store, one optimisation step and the preview run on it."""
    from PIL import Image
    from stylemesh_b200 import export, raster, synthetic as syn
    from stylemesh_b200.data.scannet_scene import ScanNetScene, load_scene_into_store
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    verts, faces, cuv, cn = rs.room_mesh()
    scene = tmp_path / "train" / "images" / "scene0000_00"
    for d in ("color", "depth", "pose"):
        os.makedirs(scene / d)
    rs.write_obj(str(tmp_path / "room.obj"), verts, faces, cuv)
    with open(scene / "scene0000_00.txt", "w") as fh:
        fh.write("colorHeight = 480\ncolorWidth = 640\nfx_color = 577.6\nfy_color = 578.7\nmx_color = 318.9\nmy_color = 242.7\n")
    g = np.random.default_rng(0)
    for i, pose in enumerate(rs.room_poses(3)):
        np.savetxt(scene / "pose" / f"{i}.txt", pose, delimiter=" ")
        Image.fromarray(g.integers(0, 255, (480, 640, 3), dtype=np.uint8)).save(scene / "color" / f"{i}.jpg")
    n = raster.render_scene(str(tmp_path / "room.obj"), str(scene / "pose"), str(scene / "scene0000_00.txt"), str(scene),
                            base_wh=(160, 120), multi_size=raster.multi_size_list(96, 192, 3, 4 / 3))
    assert n == 3 and sorted(os.listdir(scene)) == ["color", "depth", "pose", "scene0000_00.txt", "uv", "uv_144.0",
                                                    "uv_192.0", "uv_96.0"]
    sc = ScanNetScene(str(scene), pyramid_levels=3, min_pyramid_height=32)
    assert sc.rendered_depth and len(sc) == 3 and sc.levels == [96.0, 144.0, 192.0]
    store = load_scene_into_store(sc, torch.device("cuda"), 96, 0.25)
    batch = store[0]
    assert [tuple(u.shape[1:3]) for u in batch[9]] == [(96, 128), (144, 192), (192, 256)]
    assert float(batch[10].float().mean()) > 0.9                              # a closed room: nearly every pixel is valid
    preset = syn.PRESETS["with_angle_and_depth"]
    vgg_path = str(tmp_path / "vgg.pth")
    torch.save(syn.make_vgg_state_dict(0, bias_scale=0.05), vgg_path)
    mdl = TextureOptimizationStyleTransferPipeline(
        256, 256, hierarchical_texture=True, hierarchical_layers=3, random_texture_init=True,
        style_image=syn.make_style_image(7, 96, 80), style_weights=list(preset["style_weights"]),
        vgg_gatys_model_path=vgg_path, use_angle_weight=True, use_depth_scaling=True, style_pyramid_mode="multi",
        gram_mode="current", angle_threshold=60.0, learning_rate=1.0, loss_weights=dict(preset["loss_weights"]),
        save_texture=False).cuda()
    (opt,), _ = mdl.configure_optimizers()
    before = mdl.texture.layers[0].data.detach().clone()
    out = mdl.training_step(batch, 0)
    out["loss"].backward()
    opt.step()
    assert torch.isfinite(out["loss"]).all() and float(out["loss"]) > 0
    assert float((mdl.texture.layers[0].data.detach() - before).abs().max()) > 0.5   # Adam lr = 1 moved the seen texels
    uv = torch.from_numpy(np.load(sc.uv_levels[-1][0])).cuda()
    img = export.MipPreview(mdl.texture).render(uv).cpu().numpy()
    assert img.shape == (192, 256, 3) and img.std() > 1.0
