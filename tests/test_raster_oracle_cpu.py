"""The rasteriser oracle (oracle/raster_oracle.py, a restatement of the reference's OpenGL renderer) against closed-form
answers, and the host side of stylemesh_b200.raster (mesh loaders, camera matrices) - no GPU needed."""
import os
import struct

import numpy as np
import pytest

import raster_scene_util as rs
from oracle import raster_oracle as ro


def test_fronto_parallel_quad_has_closed_form_uv_depth_angle_lod():
    """A 1 m x 1 m quad at depth Z = 2 m, facing the camera, uv = position + 0.5: every covered pixel (i, j) sees
    X = (i + 0.5 - cx) Z / fx, so u = X + 0.5; depth = Z; cos(angle) = Z / |p|; one pixel spans Z / fx metres =
    1024 * Z / fx texels -> lod = log2 of that."""
    verts = np.array([[-.5, -.5, 2], [.5, -.5, 2], [.5, .5, 2], [-.5, .5, 2]], np.float32)
    faces = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    uv_v = verts[:, :2] + 0.5
    cuv = uv_v[faces]
    cn = np.tile(np.array([0, 0, -1], np.float32), (2, 3, 1))                 # facing the camera at the origin
    K = np.array([[200.0, 0, 80], [0, 200.0, 60], [0, 0, 1]])
    uv, ang, dep = ro.render(verts, faces, cuv, cn, np.eye(4), K, (160, 120), (160, 120))
    j, i = np.mgrid[0:120, 0:160]
    X, Y = (i + 0.5 - 80) * 2 / 200, (j + 0.5 - 60) * 2 / 200
    inside = (np.abs(X) < .5) & (np.abs(Y) < .5)
    assert np.array_equal(uv[..., 0] != 0, inside) or np.mean((uv[..., 0] != 0) != inside) < 1e-3
    m = inside & (uv[..., 0] != 0)
    assert m.sum() > 5000 and (~m).sum() > 5000
    assert np.allclose(uv[..., 0][m], (X + 0.5)[m], atol=1e-5) and np.allclose(uv[..., 1][m], (Y + 0.5)[m], atol=1e-5)
    assert np.allclose(dep[..., 0][m], 2.0, atol=1e-5) and np.array_equal(dep[..., 0], dep[..., 2])
    cosang = 2.0 / np.sqrt(X ** 2 + Y ** 2 + 4.0)
    assert np.allclose(ang[..., 0][m], cosang[m], atol=1e-5)
    assert np.allclose(uv[..., 2][m], np.log2(1024 * 2 / 200.0), atol=1e-4)
    assert float(uv[~m].max()) == 0 and float(dep[~m].max()) == 0           # clear colour


def test_depth_test_near_clipping_and_flip():
    verts, faces, cuv, cn = rs.room_mesh()
    pose = rs.room_poses(4)[1]
    uv, ang, dep = ro.render(verts, faces, cuv, cn, pose, rs.INTRINSICS, rs.INTRINSICS_SIZE, (128, 96))
    uvf, angf, depf = ro.render(verts, faces, cuv, cn, pose, rs.INTRINSICS, rs.INTRINSICS_SIZE, (128, 96), flip=True)
    assert np.array_equal(uvf, uv[::-1]) and np.array_equal(depf, dep[::-1]) and np.array_equal(angf, ang[::-1])
    covered = dep[..., 0] > 0
    assert covered.mean() > 0.95                                            # a closed room: geometry behind every pixel
    assert dep[covered].min() >= ro.NEAR - 1e-6 and dep[covered].max() <= 6.0
    assert ang.min() >= 0 and ang.max() <= 1 + 1e-6
    assert (uv[..., 2] >= 0).all() and (uv[..., 2] <= 10).all()
    # depth is the eye depth of the surface point: re-project a few pixels
    V = ro.view_matrix(pose)
    j, i = 40, 70
    Z = dep[j, i, 0]
    fx, fy, cx, cy = rs.INTRINSICS[0, 0], rs.INTRINSICS[1, 1], rs.INTRINSICS[0, 2], rs.INTRINSICS[1, 2]
    x_img, y_img = (i + 0.5) * 640 / 128, (j + 0.5) * 480 / 96
    p_eye = np.array([(x_img - cx) / fx * Z, (y_img - cy) / fy * Z, -Z, 1.0])
    p_world = np.linalg.inv(np.vstack([V[:3], [0, 0, 0, 1]])) @ p_eye
    lo, hi = np.array([-2.0, -1.5, -2.5]), np.array([2.0, 1.5, 2.5])
    on_wall = np.min(np.abs(np.concatenate([p_world[:3] - lo, hi - p_world[:3]]))) < 2e-3
    on_table = abs(p_world[1] + 0.35) < 0.2 and abs(p_world[0]) < 1 and abs(p_world[2]) < 1
    assert on_wall or on_table, p_world


def test_camera_matrices_of_the_product_equal_the_oracle():
    from stylemesh_b200 import raster
    pose = rs.room_poses(3)[2]
    assert np.allclose(raster.view_rows(pose), ro.view_matrix(pose)[:3])
    P = ro.projection_matrix(rs.INTRINSICS, rs.INTRINSICS_SIZE)
    e = raster.projection_entries(rs.INTRINSICS, rs.INTRINSICS_SIZE)
    assert np.allclose(e, [P[0, 0], P[0, 2], P[1, 1], P[1, 2], P[2, 2], P[2, 3]])
    assert raster.multi_size_list()[0] == (256.0, 341) and raster.multi_size_list()[-1] == (960.0, 1280)


def test_mesh_loaders_obj_and_ply(tmp_path):
    from stylemesh_b200 import raster
    verts, faces, cuv, cn = rs.room_mesh()
    obj = str(tmp_path / "room.obj")
    rs.write_obj(obj, verts, faces, cuv)
    m = raster.load_mesh(obj)
    assert np.allclose(m.verts, verts) and np.array_equal(m.faces, faces)
    assert np.allclose(m.corner_uv, cuv, atol=1e-6)                          # v flipped back like aiProcess_FlipUVs
    assert np.allclose(m.corner_normal, cn, atol=1e-6)                       # generated flat normals
    # PLY, ascii with a per-face texcoord list, and binary little endian with per-vertex s / t
    ply = str(tmp_path / "room_ascii.ply")
    with open(ply, "w") as fh:
        fh.write(f"ply\nformat ascii 1.0\nelement vertex {len(verts)}\nproperty float x\nproperty float y\nproperty float z\n"
                 f"element face {len(faces)}\nproperty list uchar int vertex_indices\nproperty list uchar float texcoord\n"
                 "end_header\n")
        for v in verts:
            fh.write(f"{v[0]:.9g} {v[1]:.9g} {v[2]:.9g}\n")
        for f, tri in enumerate(faces):
            t = " ".join(f"{cuv[f, c, 0]:.9g} {1 - cuv[f, c, 1]:.9g}" for c in range(3))
            fh.write(f"3 {tri[0]} {tri[1]} {tri[2]} 6 {t}\n")
    m2 = raster.load_mesh(ply)
    assert np.allclose(m2.verts, verts) and np.array_equal(m2.faces, faces) and np.allclose(m2.corner_uv, cuv, atol=1e-6)
    plyb = str(tmp_path / "quad_bin.ply")
    qv = np.array([[0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.float32)
    quv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    with open(plyb, "wb") as fh:
        fh.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 4\nproperty float x\nproperty float y\n"
                 b"property float z\nproperty float nx\nproperty float ny\nproperty float nz\nproperty float s\n"
                 b"property float t\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n")
        for p, t in zip(qv, quv):
            fh.write(struct.pack("<8f", *p, 0, 0, -1, *t))
        fh.write(struct.pack("<B4i", 4, 0, 1, 2, 3))                          # a quad: triangulated as a fan
    m3 = raster.load_mesh(plyb)
    assert m3.faces.tolist() == [[0, 1, 2], [0, 2, 3]] and np.allclose(m3.corner_normal, [0, 0, -1])
    assert np.allclose(m3.corner_uv[0], [[0, 1], [1, 1], [1, 0]])            # v = 1 - t
