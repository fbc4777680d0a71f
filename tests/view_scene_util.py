"""Rebuilds the synthetic ScanNet-layout scene of tests/golden/make_view_golden.py from the raw arrays stored in
tests/golden/view_prep.npz (the files the reference dataset class read when the fixture was generated)."""
import os

import numpy as np

SCENE = "scene0000_00"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "view_prep.npz")


def write_scene(gold, root) -> str:
    from PIL import Image
    n = int(gold["meta"][0])
    sp = os.path.join(str(root), "train", "images", SCENE)
    heights = sorted({int(k[2:].split("_")[0]) for k in gold.files if k.startswith("uv") and k[2].isdigit()})
    for d in ["color", "depth", "pose", "uv"] + [f"uv_{h}" for h in heights]:
        os.makedirs(os.path.join(sp, d), exist_ok=True)
    with open(os.path.join(sp, SCENE + ".txt"), "w") as f:
        f.write("colorHeight = 60\ncolorWidth = 80\nfx_color = 70.5\nfy_color = 71.25\nmx_color = 39.5\nmy_color = 29.5\n")
    for i in range(n):
        Image.fromarray(gold[f"rgb_{i}"]).save(os.path.join(sp, "color", f"{i}.png"))
        Image.fromarray(gold[f"depth_mm_{i}"]).save(os.path.join(sp, "depth", f"{i}.png"))
        np.savetxt(os.path.join(sp, "pose", f"{i}.txt"), gold[f"pose_{i}"].astype(np.float64), delimiter=" ")
        for h in heights:
            np.save(os.path.join(sp, f"uv_{h}", f"{i}.npy"), gold[f"uv{h}_{i}"])
        np.save(os.path.join(sp, "uv", f"{i}.angle.npy"), gold[f"angle_{i}"])
    return str(root)


RD_SCENE = "scene0001_00"
RD_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "view_prep_rendered.npz")


def write_rendered_depth_scene(gold, root) -> str:
    """ScanNet layout with an EMPTY depth/ folder and float32 `uv/<i>.rendered_depth.npy` files."""
    from PIL import Image
    sp = os.path.join(str(root), "train", "images", RD_SCENE)
    for d in ["color", "depth", "pose", "uv", "uv_32", "uv_48"]:
        os.makedirs(os.path.join(sp, d), exist_ok=True)
    with open(os.path.join(sp, RD_SCENE + ".txt"), "w") as f:
        f.write("colorHeight = 60\ncolorWidth = 80\nfx_color = 70.5\nfy_color = 71.25\nmx_color = 39.5\nmy_color = 29.5\n")
    for i in range(int(gold["meta"][0])):
        Image.fromarray(gold[f"rd_rgb_{i}"]).save(os.path.join(sp, "color", f"{i}.png"))
        np.savetxt(os.path.join(sp, "pose", f"{i}.txt"), np.eye(4), delimiter=" ")
        np.save(os.path.join(sp, "uv", f"{i}.rendered_depth.npy"), gold[f"rd_depth_{i}"])
        np.save(os.path.join(sp, "uv", f"{i}.angle.npy"), gold[f"rd_angle_{i}"])
        for h in (32, 48):
            np.save(os.path.join(sp, f"uv_{h}", f"{i}.npy"), gold[f"rd_uv{h}_{i}"])
    return str(root)


MP_HOUSE = "house0"
MP_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "view_prep_matterport.npz")


def write_matterport(gold, root) -> str:
    """The Matterport-layout house of make_view_golden.py from the raw arrays in view_prep_matterport.npz."""
    from PIL import Image
    rp = os.path.join(str(root), "v1", "scans", MP_HOUSE, "rendered", "region_0")
    names = [str(n) for n in gold["names"]]
    heights = sorted({int(k[5:].split("_")[0]) for k in gold.files if k.startswith("mp_uv")})
    widths = {h: gold[f"mp_uv{h}_0"].shape[1] for h in heights}
    for d in ["color", "depth", "pose", "angle"] + [f"uv_{widths[h]}_{h}" for h in heights]:
        os.makedirs(os.path.join(rp, d), exist_ok=True)
    for i in (1, 2, 0):                                  # directory order must not matter
        name = names[i]
        Image.fromarray(gold[f"mp_rgb_{i}"]).save(os.path.join(rp, "color", name + ".png"))
        Image.fromarray(gold[f"mp_depth_q_{i}"]).save(os.path.join(rp, "depth", name.replace("_i", "_d") + ".png"))
        np.savetxt(os.path.join(rp, "pose", name + ".png.pose.txt"), gold[f"mp_pose_{i}"].astype(np.float64), delimiter=" ")
        for h in heights:
            np.save(os.path.join(rp, f"uv_{widths[h]}_{h}", name + ".png.uvs.npy"), gold[f"mp_uv{h}_{i}"])
        np.save(os.path.join(rp, "angle", name + ".png.angle.npy"), gold[f"mp_angle_{i}"])
    with open(os.path.join(rp, "pose", names[0] + ".png.intrinsics.txt"), "w") as f:
        f.write("70.5 0 39.5\n0 71.25 29.5\n0 0 1\n80 60\n")
    return str(root)


NAMES = ["rgb", "extrinsics", "intrinsics", "depth", "depth_level", "rounded_depth_level", "other_depth_level",
         "interp_weight", "idx", "uv", "mask", "angle_guidance", "angle_degrees"]


def check_view_against_golden(view, gold, i):
    """view: 13-tuple of tensors with the batch dimension of default_collate (batch size 1)."""
    import torch
    assert len(view) == 13
    t = {n: v for n, v in zip(NAMES, view)}
    np_ = lambda x: x.detach().cpu().numpy()
    # integer / index / mask outputs and everything that is a pure gather or exactly rounded arithmetic: bit for bit
    assert np.array_equal(np_(t["mask"])[0], gold[f"ref_mask_{i}"]) and t["mask"].dtype == torch.bool
    assert np.array_equal(np_(t["rounded_depth_level"])[0], gold[f"ref_rounded_depth_level_{i}"])
    assert np.array_equal(np_(t["other_depth_level"])[0], gold[f"ref_other_depth_level_{i}"])
    assert t["rounded_depth_level"].dtype == torch.int64 and t["other_depth_level"].dtype == torch.int64
    nlev = int(gold["meta"][2])
    assert len(t["uv"]) == nlev
    for l in range(nlev):
        assert np.array_equal(np_(t["uv"][l])[0], gold[f"ref_uv{l}_{i}"])
    assert np.array_equal(np_(t["rgb"])[0], gold[f"ref_rgb_{i}"])
    assert np.array_equal(np_(t["angle_guidance"])[0], gold[f"ref_angle_guidance_{i}"])
    assert np.array_equal(np_(t["extrinsics"])[0], gold[f"ref_extrinsics_{i}"])
    assert np.array_equal(np_(t["intrinsics"])[0], gold[f"ref_intrinsics_{i}"])
    assert int(t["idx"][0]) == int(gold[f"ref_idx_{i}"])
    # cv2's exact-2x fast path averages the 2x2 block in one expression: <= 1 ulp of the float64 result
    assert np.allclose(np_(t["depth"])[0], gold[f"ref_depth_{i}"], rtol=1e-6, atol=0)
    assert np.allclose(np_(t["depth_level"])[0], gold[f"ref_depth_level_{i}"], rtol=0, atol=2e-6)
    assert np.allclose(np_(t["interp_weight"])[0], gold[f"ref_interp_weight_{i}"], rtol=0, atol=2e-6)
    # acos: torch (sleef) vs CUDA acosf vs numpy differ by a few float32 ulps near 90 degrees
    assert np.allclose(np_(t["angle_degrees"])[0], gold[f"ref_angle_degrees_{i}"], rtol=0, atol=3e-5)
