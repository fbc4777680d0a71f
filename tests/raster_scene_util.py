"""A small UV-parametrised test scene for the rasteriser tests: the inside of a box room (12 triangles, each wall its own
chart of a 4 x 3 UV atlas) plus a tilted table quad, cameras inside the room looking around (some triangles cross the
near plane, some leave the viewport)."""
import numpy as np


def room_mesh():
    lo, hi = np.array([-2.0, -1.5, -2.5]), np.array([2.0, 1.5, 2.5])
    c = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    verts, faces, cuv = list(c), [], []
    for qi, q in enumerate(quads):
        u0, v0 = (qi % 4) * 0.25 + 0.01, (qi // 4) * 0.33 + 0.01
        uv = np.array([[u0, v0], [u0 + 0.23, v0], [u0 + 0.23, v0 + 0.31], [u0, v0 + 0.31]])
        for tri in ((0, 1, 2), (0, 2, 3)):
            faces.append([q[t] for t in tri])
            cuv.append([uv[t] for t in tri])
    base = len(verts)                                    # a tilted table in the middle of the room
    verts += [[-0.8, -0.4, -0.6], [0.9, -0.5, -0.5], [0.8, -0.2, 0.9], [-0.7, -0.3, 0.8]]
    tuv = np.array([[0.52, 0.70], [0.98, 0.70], [0.98, 0.98], [0.52, 0.98]])
    for tri in ((0, 1, 2), (0, 2, 3)):
        faces.append([base + t for t in tri])
        cuv.append([tuv[t] for t in tri])
    verts, faces, cuv = np.asarray(verts, np.float32), np.asarray(faces, np.int32), np.asarray(cuv, np.float32)
    a, b, cc = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    n = np.cross(b - a, cc - a)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return verts, faces, cuv, np.repeat(n[:, None, :], 3, 1).astype(np.float32)


def look_at_pose(eye, target, down=(0.0, -1.0, 0.0)):
    """camera-to-world pose in the ScanNet convention: columns = camera x (right), y (DOWN), z (forward)."""
    eye, target, down = np.asarray(eye, float), np.asarray(target, float), np.asarray(down, float)
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(down, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = x, y, z, eye
    return pose


def room_poses(n=4):
    poses = []
    for i in range(n):
        a = 2 * np.pi * i / n + 0.3
        eye = np.array([0.9 * np.cos(a), 0.2 * np.sin(2 * a), 1.1 * np.sin(a)])
        tgt = np.array([-1.6 * np.cos(a + 0.5), -0.4, -2.0 * np.sin(a + 0.5)])
        poses.append(look_at_pose(eye, tgt))
    return poses


INTRINSICS = np.array([[577.6, 0, 318.9], [0, 578.7, 242.7], [0, 0, 1.0]])
INTRINSICS_SIZE = (640, 480)


def write_obj(path, verts, faces, cuv, flip_v_back=True):
    """OBJ with one vt per face corner (v is written as 1 - v: the loaders flip it back like aiProcess_FlipUVs)."""
    with open(path, "w") as fh:
        for v in verts:
            fh.write(f"v {v[0]:.9g} {v[1]:.9g} {v[2]:.9g}\n")
        for f in range(len(faces)):
            for c in range(3):
                u, v = cuv[f, c]
                fh.write(f"vt {u:.9g} {(1 - v) if flip_v_back else v:.9g}\n")
        for f, tri in enumerate(faces):
            fh.write("f " + " ".join(f"{tri[c] + 1}/{3 * f + c + 1}" for c in range(3)) + "\n")
