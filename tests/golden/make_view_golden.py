"""Golden vectors for the view-preprocessing path: a small synthetic ScanNet-layout scene is written to
a temp dir and read back through the REAL reference dataset class (`data/scannet_single_scene_dataset.py`,
`data/scannet_dataset.py:259-366`, `data/abstract_dataset.py:270-344`), imported from /root/reference with two shims
(a `pytorch_lightning` stub and `np.int`, removed from numpy >= 1.24).  The raw files' contents and the reference's
13-tuples go to tests/golden/view_prep.npz.

    python tests/golden/make_view_golden.py          (needs /root/reference, cv2 and PIL; run in the build container)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("STYLEMESH_REFERENCE", "/root/reference")

SCENE = "scene0000_00"
COLOR_HW = (60, 80)              # sensor colour / depth resolution
UV_HEIGHTS = [24, 32, 48, 64]    # uv_<h> folders; min_pyramid_height = 32 drops the first
RESIZE = 30                      # --resize_size (int => width follows the aspect ratio: 40)
MIN_DEPTH = 1.0
N_VIEWS = 3


def write_scene(root):
    from PIL import Image
    rng = np.random.default_rng(5)
    sp = os.path.join(root, "train", "images", SCENE)
    for d in ["color", "depth", "pose", "uv"] + [f"uv_{h}" for h in UV_HEIGHTS]:
        os.makedirs(os.path.join(sp, d))
    with open(os.path.join(sp, SCENE + ".txt"), "w") as f:
        f.write("colorHeight = 60\ncolorWidth = 80\nfx_color = 70.5\nfy_color = 71.25\nmx_color = 39.5\nmy_color = 29.5\n")
    raw = {}
    for i in range(N_VIEWS):
        H, W = COLOR_HW
        rgb = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        Image.fromarray(rgb).save(os.path.join(sp, "color", f"{i}.png"))
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        depth_m = 0.3 + 2.5 * ((xx / W + 0.6 * yy / H + 0.37 * i) % 1.0)          # 0.3 .. 2.8 m: spans several levels
        depth_mm = np.round(depth_m * 1000).astype(np.uint16)
        depth_mm[(xx + 2 * yy + 7 * i) % 23 == 0] = 0                             # sensor holes
        depth_mm[: 4 + i, :] = 0
        Image.fromarray(depth_mm).save(os.path.join(sp, "depth", f"{i}.png"))
        pose = np.eye(4, dtype=np.float64)
        pose[:3, 3] = rng.normal(size=3)
        np.savetxt(os.path.join(sp, "pose", f"{i}.txt"), pose, delimiter=" ")
        raw[f"rgb_{i}"] = rgb
        raw[f"depth_mm_{i}"] = depth_mm
        raw[f"pose_{i}"] = pose.astype(np.float32)
        for h in UV_HEIGHTS + [48]:
            w = h * 4 // 3
            y2, x2 = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            u = (0.1 + 0.8 * x2 / w + 0.03 * np.sin(y2 / 5.0 + i)).astype(np.float32)
            v = (0.15 + 0.7 * y2 / h + 0.02 * np.cos(x2 / 7.0 - i)).astype(np.float32)
            uv = np.stack([u, v, np.full_like(u, 1.5)], axis=2)
            hole = ((x2 * 3 // w + y2 * 2 // h + i) % 4 == 0) & (x2 > w // 3)      # contiguous invalid regions
            uv[hole] = 0.0
            if h in UV_HEIGHTS:
                np.save(os.path.join(sp, f"uv_{h}", f"{i}.npy"), uv)
                raw[f"uv{h}_{i}"] = uv
        ha, wa = 48, 64
        y3, x3 = np.meshgrid(np.arange(ha), np.arange(wa), indexing="ij")
        cosang = (0.15 + 0.85 * np.abs(np.cos(x3 / 9.0 + y3 / 13.0 + i))).astype(np.float32)
        ang = np.stack([cosang, cosang * 0, cosang * 0], axis=2)
        np.save(os.path.join(sp, "uv", f"{i}.angle.npy"), ang)
        raw[f"angle_{i}"] = ang
    return os.path.join(root, "train", "images"), raw


RD_SCENE = "scene0001_00"


def write_rendered_depth_scene(root):
    """ScanNet layout whose depth/ folder is EMPTY: the reference then reads the renderer's float32
    `uv/<i>.rendered_depth.npy` (data/scannet_dataset.py:117-144, 303-304) and keeps numpy's float32 arithmetic in
    calculate_depth_level."""
    from PIL import Image
    rng = np.random.default_rng(17)
    sp = os.path.join(root, "train", "images", RD_SCENE)
    for d in ["color", "depth", "pose", "uv", "uv_32", "uv_48"]:
        os.makedirs(os.path.join(sp, d))
    with open(os.path.join(sp, RD_SCENE + ".txt"), "w") as f:
        f.write("colorHeight = 60\ncolorWidth = 80\nfx_color = 70.5\nfy_color = 71.25\nmx_color = 39.5\nmy_color = 29.5\n")
    raw = {}
    for i in range(2):
        H, W = COLOR_HW
        rgb = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        Image.fromarray(rgb).save(os.path.join(sp, "color", f"{i}.png"))
        np.savetxt(os.path.join(sp, "pose", f"{i}.txt"), np.eye(4), delimiter=" ")
        hd, wd = 48, 64
        yy, xx = np.meshgrid(np.arange(hd), np.arange(wd), indexing="ij")
        dep = (0.25 + 2.4 * ((0.8 * xx / wd + 0.5 * yy / hd + 0.3 * i) % 1.0)).astype(np.float32)
        dep[(xx + 3 * yy + i) % 29 == 0] = 0.0
        dep3 = np.stack([dep, dep * 0, dep * 0], axis=2)
        np.save(os.path.join(sp, "uv", f"{i}.rendered_depth.npy"), dep3)
        raw[f"rd_rgb_{i}"], raw[f"rd_depth_{i}"] = rgb, dep3
        for h in (32, 48):
            w = h * 4 // 3
            y2, x2 = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            uv = np.stack([(0.1 + 0.8 * x2 / w).astype(np.float32), (0.1 + 0.8 * y2 / h).astype(np.float32),
                           np.zeros((h, w), np.float32)], axis=2)
            uv[(x2 + y2 + i) % 11 == 0] = 0.0
            np.save(os.path.join(sp, f"uv_{h}", f"{i}.npy"), uv)
            raw[f"rd_uv{h}_{i}"] = uv
        ang = np.stack([(0.2 + 0.8 * rng.random((48, 64))).astype(np.float32)] * 3, axis=2)
        np.save(os.path.join(sp, "uv", f"{i}.angle.npy"), ang)
        raw[f"rd_angle_{i}"] = ang
    return os.path.join(root, "train", "images"), raw


MP_HOUSE = "house0"
MP_NAMES = ["0e92a69a50414253_i0_0", "0e92a69a50414253_i1_2", "5b9b2794954e4694_i0_1"]   # sort: hash, camera*100 + yaw


def write_matterport(root):
    """One region of one house in the Matterport layout (data/matterport_dataset.py:96-255); reuses the ScanNet
    scene's pixel content with the Matterport file names, 1/4000 m depth units and an `.intrinsics.txt` file."""
    from PIL import Image
    rng = np.random.default_rng(9)
    rp = os.path.join(root, "v1", "scans", MP_HOUSE, "rendered", "region_0")
    sizes = {24: 32, 32: 43, 48: 64, 64: 85}
    for d in ["color", "depth", "pose", "angle"] + [f"uv_{w}_{h}" for h, w in sizes.items()]:
        os.makedirs(os.path.join(rp, d))
    raw = {}
    order = [2, 0, 1]                                  # files are written out of order: the reader has to sort them
    for i in order:
        name = MP_NAMES[i]
        H, W = COLOR_HW
        rgb = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        Image.fromarray(rgb).save(os.path.join(rp, "color", name + ".png"))
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        depth_m = 0.2 + 2.2 * ((0.7 * xx / W + yy / H + 0.21 * i) % 1.0)
        depth_q = np.round(depth_m * 4000).astype(np.uint16)
        depth_q[(3 * xx + yy + 5 * i) % 19 == 0] = 0
        Image.fromarray(depth_q).save(os.path.join(rp, "depth", name.replace("_i", "_d") + ".png"))
        pose = np.eye(4, dtype=np.float64)
        pose[:3, 3] = rng.normal(size=3)
        np.savetxt(os.path.join(rp, "pose", name + ".png.pose.txt"), pose, delimiter=" ")
        raw[f"mp_rgb_{i}"], raw[f"mp_depth_q_{i}"], raw[f"mp_pose_{i}"] = rgb, depth_q, pose.astype(np.float32)
        for h, w in sizes.items():
            y2, x2 = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            u = (0.05 + 0.9 * x2 / w + 0.02 * np.cos(y2 / 4.0 + i)).astype(np.float32)
            v = (0.1 + 0.8 * y2 / h + 0.03 * np.sin(x2 / 6.0 - i)).astype(np.float32)
            uv = np.stack([u, v, np.full_like(u, 0.5)], axis=2)
            uv[((x2 * 4 // w + y2 * 3 // h + i) % 5 == 0) & (y2 > h // 4)] = 0.0
            np.save(os.path.join(rp, f"uv_{w}_{h}", name + ".png.uvs.npy"), uv)
            raw[f"mp_uv{h}_{i}"] = uv
        ha, wa = 40, 54
        y3, x3 = np.meshgrid(np.arange(ha), np.arange(wa), indexing="ij")
        cosang = (0.1 + 0.9 * np.abs(np.sin(x3 / 8.0 - y3 / 11.0 + i))).astype(np.float32)
        ang = np.stack([cosang, cosang * 0, cosang * 0], axis=2)
        np.save(os.path.join(rp, "angle", name + ".png.angle.npy"), ang)
        raw[f"mp_angle_{i}"] = ang
    with open(os.path.join(rp, "pose", MP_NAMES[0] + ".png.intrinsics.txt"), "w") as f:
        f.write("70.5 0 39.5\n0 71.25 29.5\n0 0 1\n80 60\n")
    return os.path.join(root, "v1", "scans"), raw


def import_reference():
    if not hasattr(np, "int"):
        np.int = int                                       # scannet_dataset.py:365
    pl = types.ModuleType("pytorch_lightning")

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    class LightningModule(torch.nn.Module):
        pass

    pl.LightningDataModule, pl.LightningModule = LightningDataModule, LightningModule
    sys.modules["pytorch_lightning"] = pl
    sys.path.insert(0, REF)
    from data.scannet_single_scene_dataset import ScanNet_Single_House_Dataset
    from data.matterport_single_scene_dataset import Matterport_Single_House_Dataset
    from model.texture.utils import get_rgb_transform, get_label_transform, get_uv_transform
    from model.losses.rgb_transform import pre
    from torchvision.transforms import Compose
    return (ScanNet_Single_House_Dataset, Matterport_Single_House_Dataset, Compose([get_rgb_transform(), pre()]),
            get_label_transform(), get_uv_transform())


def main():
    tmp = tempfile.mkdtemp(prefix="smb_view_golden_")
    root, raw = write_scene(tmp)
    DS, MPDS, t_rgb, t_label, t_uv = import_reference()
    ds = DS(root_path=root, scene=SCENE, min_images=1, max_images=-1, transform_rgb=t_rgb, transform_label=t_label,
            transform_uv=t_uv, resize_size=RESIZE, pyramid_levels=3, min_pyramid_depth=MIN_DEPTH,
            min_pyramid_height=32, verbose=False)
    out = dict(raw)
    out["levels"] = np.asarray(ds.levels, dtype=np.float64)
    out["all_levels"] = np.asarray(ds.all_levels, dtype=np.float64)
    names = ["rgb", "extrinsics", "intrinsics", "depth", "depth_level", "rounded_depth_level", "other_depth_level",
             "interp_weight", "idx", "uv", "mask", "angle_guidance", "angle_degrees"]
    for i in range(len(ds)):
        item = ds[i]
        assert len(item) == 13
        for n, t in zip(names, item):
            if n == "uv":
                for l, u in enumerate(t):
                    out[f"ref_uv{l}_{i}"] = np.asarray(u)
            elif n == "idx":
                out[f"ref_idx_{i}"] = np.asarray(t)
            else:
                out[f"ref_{n}_{i}"] = np.asarray(t)
    out["meta"] = np.asarray([N_VIEWS, RESIZE, len(ds.levels)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "view_prep.npz"), **out)

    # rendered (float32) depth instead of sensor PNGs
    rd_tmp = tempfile.mkdtemp(prefix="smb_view_golden_rd_")
    rd_root, rd_raw = write_rendered_depth_scene(rd_tmp)
    rd = DS(root_path=rd_root, scene=RD_SCENE, min_images=1, max_images=-1, transform_rgb=t_rgb, transform_label=t_label,
            transform_uv=t_uv, resize_size=RESIZE, pyramid_levels=2, min_pyramid_depth=MIN_DEPTH, min_pyramid_height=32,
            verbose=False)
    assert rd.rendered_depth
    ro = dict(rd_raw)
    ro["levels"] = np.asarray(rd.levels, dtype=np.float64)
    for i in range(len(rd)):
        item = rd[i]
        for n, t in zip(names, item):
            if n == "uv":
                for l, u in enumerate(t):
                    ro[f"ref_uv{l}_{i}"] = np.asarray(u)
            elif n == "idx":
                ro[f"ref_idx_{i}"] = np.asarray(t)
            else:
                ro[f"ref_{n}_{i}"] = np.asarray(t)
    ro["meta"] = np.asarray([len(rd), RESIZE, len(rd.levels)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "view_prep_rendered.npz"), **ro)
    print("rendered-depth levels", ro["levels"], "views", len(rd), "depth dtype", ro["ref_depth_0"].dtype)

    # Matterport: same pipeline, other file layout, depth / 4000, mask without the depth factor
    mp_root, mp_raw = write_matterport(tmp)
    mp = MPDS(root_path=mp_root, scene=MP_HOUSE, min_images=1, max_images=-1, transform_rgb=t_rgb,
              transform_label=t_label, transform_uv=t_uv, resize_size=RESIZE, pyramid_levels=3,
              min_pyramid_depth=MIN_DEPTH, min_pyramid_height=32, region_index=0, verbose=False)
    mo = dict(mp_raw)
    mo["levels"] = np.asarray(mp.levels, dtype=np.float64)
    mo["all_levels"] = np.asarray(mp.all_levels, dtype=np.float64)
    mo["names"] = np.asarray(MP_NAMES)
    for i in range(len(mp)):
        item = mp[i]
        for n, t in zip(names, item):
            if n == "uv":
                for l, u in enumerate(t):
                    mo[f"ref_uv{l}_{i}"] = np.asarray(u)
            elif n == "idx":
                mo[f"ref_idx_{i}"] = np.asarray(t)
            else:
                mo[f"ref_{n}_{i}"] = np.asarray(t)
    mo["meta"] = np.asarray([len(mp), RESIZE, len(mp.levels)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "view_prep_matterport.npz"), **mo)
    print("matterport levels", mo["levels"], "views", len(mp), "bytes",
          os.path.getsize(os.path.join(HERE, "view_prep_matterport.npz")))
    for k in sorted(out):
        if k.startswith("ref_") and k.endswith("_0"):
            print(k, out[k].dtype, out[k].shape)
    print("levels", out["levels"], "bytes", os.path.getsize(os.path.join(HERE, "view_prep.npz")))


if __name__ == "__main__":
    main()
