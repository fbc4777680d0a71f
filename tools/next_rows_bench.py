"""Timing of the kernels of SURVEY §8(f) rows 3 and 4 and of K5 on a B200, each with the CPU chain it replaces timed
beside it (the oracle restatement / the reference's torch ops on the host cores).  Not a test; prints one JSON line per
row.  CUDA events on the launching stream, 3 warm-up + 20 timed repetitions.

    python tools/next_rows_bench.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from stylemesh_b200 import export, raster, synthetic as syn                       # noqa: E402
from stylemesh_b200.model.texture.texture import HierarchicalNeuralTexture        # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def cuda_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def host_s(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def tessellated_room(k):
    """tests/raster_scene_util.room_mesh with every triangle split into 4**k (positions and UVs are linear per face)."""
    from raster_scene_util import room_mesh
    verts, faces, cuv, cn = room_mesh()
    P = verts[faces]                                   # (F, 3, 3)
    for _ in range(k):
        def split(A):
            a, b, c = A[:, 0], A[:, 1], A[:, 2]
            ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
            return np.concatenate([np.stack(t, 1) for t in ((a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca))], 0)
        P, cuv, cn = split(P), split(cuv), split(cn)
    F = P.shape[0]
    return raster.Mesh(np.ascontiguousarray(P.reshape(-1, 3), np.float32), np.arange(3 * F, dtype=np.int32).reshape(F, 3),
                       np.ascontiguousarray(cuv, np.float32), np.ascontiguousarray(cn, np.float32))


def bench_raster(quick):
    from raster_scene_util import INTRINSICS, INTRINSICS_SIZE, room_poses
    from oracle import raster_oracle as ro
    poses = room_poses(8)
    rows = []
    for k in ([3, 6] if quick else [3, 6, 7, 8]):
        mesh = tessellated_room(k)
        r = raster.MeshRasterizer(mesh)
        for wh in ((640, 480), (1045, 784)):
            i = [0]

            def one():
                r.render(poses[i[0] % len(poses)], INTRINSICS, INTRINSICS_SIZE, wh)
                i[0] += 1
            ms = cuda_ms(one, reps=16)
            rows.append({"faces": int(mesh.faces.shape[0]), "size_wh": list(wh), "ms_per_pose": round(ms, 4),
                         "poses_per_s": round(1e3 / ms, 1),
                         "out_bytes": 3 * 3 * 4 * wh[0] * wh[1]})
    # CPU restatement (numpy float64, one face at a time) on the un-tessellated room: the only CPU renderer there is here
    from raster_scene_util import room_mesh
    v, f, cuv, cn = room_mesh()
    cpu = host_s(lambda: ro.render(v, f, cuv, cn, poses[0], INTRINSICS, INTRINSICS_SIZE, (640, 480)))
    return {"row": "f4 rasteriser (smb_raster_view: vertex + z-buffer + resolve)", "gpu": rows,
            "cpu_oracle": {"faces": int(f.shape[0]), "size_wh": [640, 480], "s_per_pose": round(cpu, 3),
                           "kind": "oracle/raster_oracle.py (numpy float64; the reference renderer is OpenGL and "
                                   "cannot run without a GL context)"},
            "bound": "latency / atomics: one 64-bit atomicMin per covered pixel and face, three (h, w, 3) fp32 maps out"}


def reference_export_chain(layers):
    """texture.py:110-121 get_image + rgb_transform.py:14-21 post() + texture.py:9-19 ToPILImage on the host cores."""
    import torch.nn.functional as F
    C, H, W = layers[0].shape
    w_range = torch.arange(0, W, dtype=torch.float) / (W - 1.0) * 2.0 - 1.0
    h_range = torch.arange(0, H, dtype=torch.float) / (H - 1.0) * 2.0 - 1.0
    v, u = torch.meshgrid(h_range, w_range, indexing="ij")
    grid = torch.stack([u, v], 2).unsqueeze(0)
    img = sum(F.grid_sample(l.clamp(-123.68, 151.061).unsqueeze(0), grid, mode="bilinear", padding_mode="border",
                            align_corners=True) for l in layers)[0]
    x = img.mul(1.0 / 255)
    x = x - torch.tensor([-0.40760392, -0.45795686, -0.48501961]).view(3, 1, 1)
    x = x[torch.LongTensor([2, 1, 0])].clamp(0, 1)
    return x.mul(255).byte().permute(1, 2, 0).contiguous()


def bench_export(quick):
    hbm = float(peaks().get("hbm_gbs", 6400.0))
    rows = []
    for T in ([2048] if quick else [2048, 4096]):
        g = torch.Generator().manual_seed(T)
        ls = [(torch.rand(3, T >> i, T >> i, generator=g) * 300 - 140) / (i + 1) for i in range(4)]
        tex = HierarchicalNeuralTexture.from_tensor([l.clone() for l in ls]).cuda()
        ms = cuda_ms(lambda: export.texture_rgb8(tex), reps=10)
        ms_d2h = cuda_ms(lambda: export.texture_rgb8(tex).cpu(), reps=5)
        # algorithmic bytes: every layer read once, the composed fp32 image written and read once, 3 bytes per texel out
        alg = sum(l.numel() for l in ls) * 4 + 2 * 3 * T * T * 4 + 3 * T * T
        cpu = host_s(lambda: reference_export_chain(ls))
        rows.append({"texture": T, "layers": 4, "gpu_ms": round(ms, 4), "gpu_ms_with_d2h_of_the_bytes": round(ms_d2h, 3),
                     "algorithmic_bytes": alg, "gbs": round(alg / ms / 1e6, 1), "frac_of_hbm_peak": round(alg / ms / 1e6 / hbm, 3),
                     "cpu_chain_s": round(cpu, 3), "cpu_threads": torch.get_num_threads()})
        mips = export.MipPreview(tex)
        uv = torch.rand(480, 640, 3, generator=g)
        uv[..., 2] *= 6.0
        uvd = uv.cuda()
        rows[-1]["mip_chain_build_ms"] = round(cuda_ms(lambda: export.MipPreview(tex), reps=5), 4)
        rows[-1]["mip_preview_640x480_ms"] = round(cuda_ms(lambda: mips.render(uvd), reps=20), 4)
        del tex, mips
    return {"row": "f3 texture export (get_image -> post() -> bytes) and mip-mapped preview", "gpu": rows,
            "bound": "hbm", "hbm_peak_gbs": hbm}


def bench_k5(quick):
    """First-visit cost of a view's plan (K5 kernels + one read-back) for the 4-level with_angle_and_depth family."""
    import tempfile
    from oracle import stylemesh_oracle as orc
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    preset = syn.PRESETS["with_angle_and_depth"]
    tmp = tempfile.mkdtemp()
    vgg_path = os.path.join(tmp, "vgg.pth")
    torch.save(syn.make_vgg_state_dict(0, bias_scale=0.0), vgg_path)
    mdl = TextureOptimizationStyleTransferPipeline(
        256, 256, hierarchical_texture=True, hierarchical_layers=4, random_texture_init=True,
        style_image=syn.make_style_image(7, 96, 80), style_weights=list(preset["style_weights"]),
        vgg_gatys_model_path=vgg_path, use_angle_weight=preset["use_angle_weight"],
        use_depth_scaling=preset["use_depth_scaling"], style_pyramid_mode=preset["style_pyramid_mode"],
        gram_mode=preset["gram_mode"], angle_threshold=preset["angle_threshold"], learning_rate=1.0,
        loss_weights=dict(preset["loss_weights"]), save_texture=False).cuda()
    rows = []
    for name, table in (("scannet", syn.SCANNET_PYRAMID), ("matterport", syn.MATTERPORT_PYRAMID)):
        sizes = list(table)
        view = syn.make_view(1000, sizes[0], sizes)
        dev = view.to(torch.device("cuda", 0)).as_batch()
        mdl.build_view_plan(dev)                                       # module load, allocator warm-up
        torch.cuda.synchronize()
        gpu = host_s(lambda: (mdl.build_view_plan(dev), torch.cuda.synchronize()), reps=20)
        host = view.as_batch()
        cpu = host_s(lambda: orc.level_masks_and_weights(host, sizes, True), reps=3)
        rows.append({"pyramid": name, "levels": len(sizes), "gpu_ms_per_view_wall": round(gpu * 1e3, 3),
                     "cpu_level_masks_only_ms": round(cpu * 1e3, 2)})
    return {"row": "K5 view plan: view_level_masks + view_level_plan + ONE read-back of the counts, first visit of a view "
                   "(host wall clock, launch overhead included); CPU leg = only the level masks / weights of "
                   "model.py:210-239 in the oracle (the per-layer masks of cs:161-185 are extra there)", "gpu": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--rows", default="f3,f4,k5")
    a = ap.parse_args()
    assert torch.cuda.is_available(), "needs a GPU"
    for name, fn in (("f4", bench_raster), ("f3", bench_export), ("k5", bench_k5)):
        if name in a.rows.split(","):
            print(json.dumps(fn(a.quick)), flush=True)
