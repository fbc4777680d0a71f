"""CPU oracle of the reference's UV / angle / depth renderer.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy float64, what `scripts/scannet/render_uv` (C++17 + OpenGL 3.3/4.0) draws for one pose:

  * camera: `Scannet_Renderer::renderTrajectory` (src/renderer/scannet_renderer.cpp:19-62) builds the view matrix from
    the camera-to-world pose file ([R1 R2 R3 | T] = right, "up" (the camera's y axis, which points DOWN in ScanNet), look)
    as rows (right, up, -look); `camera_utils::perspective` (include/util.h:11-35) builds the projection from the
    intrinsics and the size they refer to, near 0.1, far 10;
  * vertex stage: shader/uvmap.vs, angle.vs, depth.vs - gl_Position = P V M p (M = identity), eye-space normal
    (inverse-transpose of V, a rotation) and eye-space position for the angle pass;
  * fixed function: clipping against the near plane, perspective division, viewport transform to w x h pixels, pixel
    centres at (i + 0.5, j + 0.5), top-left fill rule, perspective-correct attribute interpolation, depth test LESS
    (the first triangle drawn wins a tie), no face culling, clear colour 0;
  * fragment stage: uvmap.frag:8-13 writes (u, v, textureQueryLod(tex, uv).x) with the 1024 x 1024 dummy texture of
    renderer.cpp:121-139 (GL_LINEAR_MIPMAP_LINEAR, no anisotropy): lod = clamp(log2(rho), 0, 10) with rho the larger of
    the two screen-space derivative lengths of (1024 u, 1024 v); angle.frag:22-33 writes max(dot(normalize(n),
    normalize(-fragPos)), 0) three times; depth.frag:11-19 writes LinearizeDepth(gl_FragCoord.z) = the eye depth;
  * read-back: `Renderer::saveUV` (renderer.cpp:197-224): GL row j of the framebuffer becomes row j of the (h, w, 3)
    float32 .npy, or row h-1-j with flip.  With the "up" row of the view matrix being the camera's DOWN axis, GL row j IS
    image row j of the camera, so flip = 0 gives upright maps (render_uvs.py:42-45 flips only the hand-made trajectories).

Parity pinning: **unpinned** - the reference renderer needs an OpenGL context, GLFW, GLEW, Assimp and OpenCV, none of
which exist here, and it ships no rendered fixtures.  Derivatives: GL hardware differences attributes inside 2x2 pixel
quads; this oracle (and the CUDA kernel) use the analytic derivative of the perspective-correct interpolant, which is
what the quad differences approximate.  `Assimp aiProcess_FlipUVs` (model.h:57): v is stored as 1 - v_file; the
loaders in stylemesh_b200.raster do the same.
"""
from __future__ import annotations

import numpy as np

NEAR, FAR = 0.1, 10.0          # scannet_renderer.h kNearPlane / kFarPlane, angle.frag / depth.frag near / far
TEX_SIZE = 1024                # renderer.cpp:121 dummy texture the LOD is queried against


def view_matrix(pose_c2w: np.ndarray) -> np.ndarray:
    """scannet_renderer.cpp:24-55."""
    pose = np.asarray(pose_c2w, dtype=np.float64)
    right, up, look, eye = pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3]
    right, up, look = right / np.linalg.norm(right), up / np.linalg.norm(up), look / np.linalg.norm(look)
    V = np.eye(4)
    V[0, :3], V[0, 3] = right, -right @ eye
    V[1, :3], V[1, 3] = up, -up @ eye
    V[2, :3], V[2, 3] = -look, look @ eye
    return V


def projection_matrix(K: np.ndarray, size_wh, n: float = NEAR, f: float = FAR) -> np.ndarray:
    """include/util.h:11-35 (row-major as written there)."""
    W, H = size_wh
    return np.array([[2 * K[0, 0] / W, 0, -(2 * (K[0, 2] / W) - 1), 0],
                     [0, 2 * K[1, 1] / H, -(2 * (K[1, 2] / H) - 1), 0],
                     [0, 0, -(f + n) / (f - n), -2 * f * n / (f - n)],
                     [0, 0, -1, 0]], dtype=np.float64)


def _clip_near(poly):
    """Sutherland-Hodgman against z_c >= -w_c on a list of (clip position (4,), attribute vector)."""
    out = []
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        da, db = a[0][2] + a[0][3], b[0][2] + b[0][3]
        if da >= 0:
            out.append(a)
        if (da >= 0) != (db >= 0):
            t = da / (da - db)
            out.append((a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1])))
    return out


def render(verts, faces, corner_uv, corner_n, pose_c2w, K, K_size_wh, out_wh, flip=False, tex_size=TEX_SIZE,
           near=NEAR, far=FAR):
    """verts (V,3), faces (F,3) int, corner_uv (F,3,2), corner_n (F,3,3) -> uv (h,w,3), angle (h,w,3), depth (h,w,3)
    float32, like the three .npy files main.cpp:60-67 writes per pose."""
    w, h = out_wh
    V, P = view_matrix(pose_c2w), projection_matrix(K, K_size_wh, near, far)
    R = V[:3, :3]
    pe = (V @ np.concatenate([np.asarray(verts, np.float64), np.ones((len(verts), 1))], 1).T).T       # eye space
    zbuf = np.full((h, w), np.inf)
    uv_o, ang_o, dep_o = np.zeros((h, w, 3)), np.zeros((h, w, 3)), np.zeros((h, w, 3))
    for f_i, tri in enumerate(np.asarray(faces)):
        poly = []
        for c in range(3):
            p_eye = pe[tri[c]]
            attr = np.concatenate([corner_uv[f_i, c], R @ corner_n[f_i, c], p_eye[:3]])               # u v | n | fragPos
            poly.append((P @ p_eye, attr.astype(np.float64)))
        poly = _clip_near(poly)
        for k in range(1, len(poly) - 1):
            tri_c = [poly[0], poly[k], poly[k + 1]]
            wc = np.array([t[0][3] for t in tri_c])
            ndc = np.array([t[0][:3] / t[0][3] for t in tri_c])
            sx, sy = (ndc[:, 0] + 1) * 0.5 * w, (ndc[:, 1] + 1) * 0.5 * h
            zw = (ndc[:, 2] + 1) * 0.5
            area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sx[2] - sx[0]) * (sy[1] - sy[0])
            if area == 0:
                continue
            x0, x1 = max(int(np.floor(sx.min() - 0.5)), 0), min(int(np.ceil(sx.max() - 0.5)), w - 1)
            y0, y1 = max(int(np.floor(sy.min() - 0.5)), 0), min(int(np.ceil(sy.max() - 0.5)), h - 1)
            if x0 > x1 or y0 > y1:
                continue
            px, py = np.meshgrid(np.arange(x0, x1 + 1) + 0.5, np.arange(y0, y1 + 1) + 0.5)
            lam, inside = [], np.ones_like(px, dtype=bool)
            for i in range(3):
                a, b = (i + 1) % 3, (i + 2) % 3
                ex, ey = sx[b] - sx[a], sy[b] - sy[a]
                e = (ex * (py - sy[a]) - ey * (px - sx[a])) * np.sign(area)          # >= 0 inside (any winding)
                exs, eys = ex * np.sign(area), ey * np.sign(area)
                top_left = (eys < 0) or (eys == 0 and exs > 0)                       # y grows with the GL row index
                inside &= (e > 0) | ((e == 0) & top_left)
                lam.append(e / abs(area))
            if not inside.any():
                continue
            lam = np.stack(lam)                                                       # screen-space barycentrics
            z_win = np.tensordot(zw, lam, 1)
            persp = lam / wc[:, None, None]
            denom = persp.sum(0)
            attrs = np.stack([t[1] for t in tri_c])                                   # (3, 8)
            val = np.tensordot(attrs.T, persp, 1) / denom                             # (8, rows, cols)
            # analytic screen-space derivatives of u, v (quotient rule on N / D, both affine in x, y)
            dlam = np.zeros((3, 2))
            for i in range(3):
                a, b = (i + 1) % 3, (i + 2) % 3
                dlam[i] = np.array([-(sy[b] - sy[a]), (sx[b] - sx[a])]) * np.sign(area) / abs(area)
            dD = (dlam / wc[:, None]).sum(0)
            rho = np.zeros_like(denom)
            grads = []
            for ch in range(2):
                dN = (dlam * (attrs[:, ch] / wc)[:, None]).sum(0)
                gx = (dN[0] - val[ch] * dD[0]) / denom
                gy = (dN[1] - val[ch] * dD[1]) / denom
                grads.append((gx * tex_size, gy * tex_size))
            rho = np.maximum(np.sqrt(grads[0][0] ** 2 + grads[1][0] ** 2), np.sqrt(grads[0][1] ** 2 + grads[1][1] ** 2))
            lod = np.clip(np.log2(np.maximum(rho, 1e-30)), 0.0, np.log2(tex_size))
            n = val[2:5]
            n = n / np.maximum(np.linalg.norm(n, axis=0), 1e-30)
            vdir = -val[5:8]
            vdir = vdir / np.maximum(np.linalg.norm(vdir, axis=0), 1e-30)
            diff = np.maximum((n * vdir).sum(0), 0.0)
            z_ndc = z_win * 2 - 1
            lin = (2 * near * far) / (far + near - z_ndc * (far - near))
            rows, cols = np.arange(y0, y1 + 1), np.arange(x0, x1 + 1)
            sub = zbuf[np.ix_(rows, cols)]
            win = inside & (z_win >= 0) & (z_win <= 1) & (z_win < sub)
            if not win.any():
                continue
            rr, cc = np.nonzero(win)
            zbuf[rows[rr], cols[cc]] = z_win[rr, cc]
            uv_o[rows[rr], cols[cc]] = np.stack([val[0][rr, cc], val[1][rr, cc], lod[rr, cc]], -1)
            ang_o[rows[rr], cols[cc]] = diff[rr, cc][:, None]
            dep_o[rows[rr], cols[cc]] = lin[rr, cc][:, None]
    if flip:
        uv_o, ang_o, dep_o = uv_o[::-1], ang_o[::-1], dep_o[::-1]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return f32(uv_o), f32(ang_o), f32(dep_o)
