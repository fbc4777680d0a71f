// UV / angle / depth rasteriser: what the reference's stand-alone OpenGL renderer
// (scripts/scannet/render_uv, scripts/matterport/render_uv; C++17 + OpenGL 3.3/4.0) draws for one camera pose, as three
// CUDA kernels - no GL context, no window, output straight into device tensors in the layout of its .npy files:
//
//   uv    (h, w, 3) = (u, v, textureQueryLod(tex1024, uv).x)                      shader/uvmap.frag:8-13
//   angle (h, w, 3) = max(dot(normalize(n_eye), normalize(-p_eye)), 0) x 3        shader/angle.vs, angle.frag:22-33
//   depth (h, w, 3) = LinearizeDepth(gl_FragCoord.z) x 3 = eye depth              shader/depth.frag:11-19
//
// Pipeline (restated step by step in oracle/raster_oracle.py, which also lists the reference lines):
//   raster_vertex    eye-space position of every vertex (view matrix of scannet_renderer.cpp:24-55)
//   raster_zbuffer   one WARP per face: clip-space setup (projection of include/util.h:11-35), near-plane clipping
//                    (up to two triangles), viewport transform, lanes stride over the bounding box, pixel centres at
//                    +0.5, top-left rule, window depth by screen-space interpolation, atomicMin of (depth bits, face id)
//                    into a 64-bit z-buffer: depth test LESS, the lower face index wins a tie like the earlier draw
//   raster_resolve   one thread per pixel: re-derive the winning (sub)triangle, perspective-correct attributes,
//                    analytic screen-space derivatives of (u, v) for the LOD, the three outputs (optionally row-flipped,
//                    Renderer::saveUV renderer.cpp:197-224); pixels without geometry keep the clear colour 0
// Written for exactness against the oracle and for simplicity - it runs once per pose and pyramid size (a few hundred
// thousand small triangles), not on the optimisation hot path.
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

struct RasterCam {
  float V[12];                       // rows (right | -right.eye), (up | -up.eye), (-look | look.eye)
  float p00, p02, p11, p12, p22, p23;
  int w, h, flip;
  float near, far, tex_size;
};

struct ClipVert {
  float c[4];                        // clip-space position
  float a[8];                        // u v | eye normal | eye position
};

__device__ __forceinline__ void lerp_vert(const ClipVert& a, const ClipVert& b, float t, ClipVert& o, int na) {
#pragma unroll
  for (int i = 0; i < 4; ++i) o.c[i] = a.c[i] + t * (b.c[i] - a.c[i]);
  for (int i = 0; i < na; ++i) o.a[i] = a.a[i] + t * (b.a[i] - a.a[i]);
}

// Sutherland-Hodgman against z_c >= -w_c; returns the vertex count of the clipped polygon (0, 3 or 4)
__device__ __forceinline__ int clip_near(const ClipVert (&in)[3], ClipVert (&out)[4], int na) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const ClipVert& a = in[i];
    const ClipVert& b = in[(i + 1) % 3];
    const float da = a.c[2] + a.c[3], db = b.c[2] + b.c[3];
    if (da >= 0.f) out[n++] = a;
    if ((da >= 0.f) != (db >= 0.f)) lerp_vert(a, b, da / (da - db), out[n++], na);
  }
  return n;
}

struct ScreenTri {
  float sx[3], sy[3], zw[3], wc[3];
  float area;                         // signed; sgn = its sign
};

__device__ __forceinline__ bool screen_setup(const ClipVert& v0, const ClipVert& v1, const ClipVert& v2,
                                             const RasterCam& cam, ScreenTri& t) {
  const ClipVert* v[3] = {&v0, &v1, &v2};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float iw = 1.f / v[i]->c[3];
    t.wc[i] = v[i]->c[3];
    t.sx[i] = (v[i]->c[0] * iw + 1.f) * 0.5f * cam.w;
    t.sy[i] = (v[i]->c[1] * iw + 1.f) * 0.5f * cam.h;
    t.zw[i] = (v[i]->c[2] * iw + 1.f) * 0.5f;
  }
  t.area = (t.sx[1] - t.sx[0]) * (t.sy[2] - t.sy[0]) - (t.sx[2] - t.sx[0]) * (t.sy[1] - t.sy[0]);
  return t.area != 0.f && isfinite(t.area);
}

// screen-space barycentrics of the pixel centre (px, py) and the coverage test (top-left rule as in the oracle)
__device__ __forceinline__ bool cover(const ScreenTri& t, float px, float py, float (&lam)[3]) {
  const float sgn = t.area > 0.f ? 1.f : -1.f, inv = 1.f / fabsf(t.area);
  bool inside = true;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int a = (i + 1) % 3, b = (i + 2) % 3;
    const float ex = (t.sx[b] - t.sx[a]) * sgn, ey = (t.sy[b] - t.sy[a]) * sgn;
    const float e = ex * (py - t.sy[a]) - ey * (px - t.sx[a]);
    const bool top_left = (ey < 0.f) || (ey == 0.f && ex > 0.f);
    inside = inside && (e > 0.f || (e == 0.f && top_left));
    lam[i] = e * inv;
  }
  return inside;
}

__device__ __forceinline__ void load_face(const float4* __restrict__ eye, const int* __restrict__ faces,
                                          const float* __restrict__ cuv, const float* __restrict__ cn, int f,
                                          const RasterCam& cam, ClipVert (&v)[3], bool attrs) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 p = eye[faces[3 * f + c]];
    v[c].c[0] = cam.p00 * p.x + cam.p02 * p.z;
    v[c].c[1] = cam.p11 * p.y + cam.p12 * p.z;
    v[c].c[2] = cam.p22 * p.z + cam.p23;
    v[c].c[3] = -p.z;
    if (attrs) {
      v[c].a[0] = cuv[(3 * f + c) * 2];
      v[c].a[1] = cuv[(3 * f + c) * 2 + 1];
      const float nx = cn[(3 * f + c) * 3], ny = cn[(3 * f + c) * 3 + 1], nz = cn[(3 * f + c) * 3 + 2];
      v[c].a[2] = cam.V[0] * nx + cam.V[1] * ny + cam.V[2] * nz;       // rotation part of the view matrix
      v[c].a[3] = cam.V[4] * nx + cam.V[5] * ny + cam.V[6] * nz;
      v[c].a[4] = cam.V[8] * nx + cam.V[9] * ny + cam.V[10] * nz;
      v[c].a[5] = p.x; v[c].a[6] = p.y; v[c].a[7] = p.z;
    }
  }
}

__global__ void __launch_bounds__(256) raster_vertex_kernel(const float* __restrict__ verts, int nv, const RasterCam cam,
                                                            float4* __restrict__ eye) {
  pdl_sync();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    eye[i] = make_float4(cam.V[0] * x + cam.V[1] * y + cam.V[2] * z + cam.V[3],
                         cam.V[4] * x + cam.V[5] * y + cam.V[6] * z + cam.V[7],
                         cam.V[8] * x + cam.V[9] * y + cam.V[10] * z + cam.V[11], 1.f);
  }
}

__global__ void __launch_bounds__(256) raster_zbuffer_kernel(const float4* __restrict__ eye, const int* __restrict__ faces,
                                                             int nf, const RasterCam cam,
                                                             unsigned long long* __restrict__ zbuf) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < nf; f += warps) {
    ClipVert v[3], poly[4];
    load_face(eye, faces, nullptr, nullptr, f, cam, v, false);
    const int n = clip_near(v, poly, 0);
    for (int k = 0; k + 2 < n; ++k) {
      ScreenTri t;
      if (!screen_setup(poly[0], poly[k + 1], poly[k + 2], cam, t)) continue;
      const float xmin = fminf(t.sx[0], fminf(t.sx[1], t.sx[2])), xmax = fmaxf(t.sx[0], fmaxf(t.sx[1], t.sx[2]));
      const float ymin = fminf(t.sy[0], fminf(t.sy[1], t.sy[2])), ymax = fmaxf(t.sy[0], fmaxf(t.sy[1], t.sy[2]));
      if (!(xmax >= 0.f && ymax >= 0.f && xmin <= (float)cam.w && ymin <= (float)cam.h)) continue;   // also drops NaN
      const int x0 = max((int)floorf(xmin - 0.5f), 0), x1 = min((int)ceilf(xmax - 0.5f), cam.w - 1);
      const int y0 = max((int)floorf(ymin - 0.5f), 0), y1 = min((int)ceilf(ymax - 0.5f), cam.h - 1);
      if (x0 > x1 || y0 > y1) continue;
      const int bw = x1 - x0 + 1;
      const long long npx = (long long)bw * (y1 - y0 + 1);
      for (long long i = lane; i < npx; i += 32) {
        const int x = x0 + (int)(i % bw), y = y0 + (int)(i / bw);
        float lam[3];
        if (!cover(t, x + 0.5f, y + 0.5f, lam)) continue;
        const float z = lam[0] * t.zw[0] + lam[1] * t.zw[1] + lam[2] * t.zw[2];
        if (!(z >= 0.f && z <= 1.f)) continue;
        const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned)(2 * f + k);
        atomicMin(zbuf + (long long)y * cam.w + x, key);
      }
    }
  }
}

__global__ void __launch_bounds__(256) raster_resolve_kernel(const float4* __restrict__ eye, const int* __restrict__ faces,
                                                             const float* __restrict__ cuv, const float* __restrict__ cn,
                                                             const RasterCam cam,
                                                             const unsigned long long* __restrict__ zbuf,
                                                             float* __restrict__ uv_out, float* __restrict__ ang_out,
                                                             float* __restrict__ dep_out) {
  pdl_sync();
  const long long n = (long long)cam.w * cam.h, stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const int x = (int)(p % cam.w), y = (int)(p / cam.w);
    const long long o = ((long long)(cam.flip ? cam.h - 1 - y : y) * cam.w + x) * 3;
    float r_uv[3] = {0.f, 0.f, 0.f}, r_ang = 0.f, r_dep = 0.f;
    const unsigned long long key = zbuf[p];
    if (key != ~0ull) {
      const unsigned id = (unsigned)(key & 0xffffffffu);
      const int f = (int)(id >> 1), k = (int)(id & 1u);
      ClipVert v[3], poly[4];
      load_face(eye, faces, cuv, cn, f, cam, v, true);
      clip_near(v, poly, 8);
      ScreenTri t;
      screen_setup(poly[0], poly[k + 1], poly[k + 2], cam, t);
      const ClipVert* tv[3] = {&poly[0], &poly[k + 1], &poly[k + 2]};
      float lam[3];
      cover(t, x + 0.5f, y + 0.5f, lam);
      float pw[3], denom = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        pw[i] = lam[i] / t.wc[i];
        denom += pw[i];
      }
      float val[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) val[c] = (pw[0] * tv[0]->a[c] + pw[1] * tv[1]->a[c] + pw[2] * tv[2]->a[c]) / denom;
      // analytic derivatives of u, v: both numerator and denominator of the interpolant are affine in (x, y)
      const float sgn = t.area > 0.f ? 1.f : -1.f, inv = 1.f / fabsf(t.area);
      float dDx = 0.f, dDy = 0.f, dNx[2] = {0.f, 0.f}, dNy[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        const float lx = -(t.sy[b] - t.sy[a]) * sgn * inv, ly = (t.sx[b] - t.sx[a]) * sgn * inv;
        dDx += lx / t.wc[i];
        dDy += ly / t.wc[i];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          dNx[c] += lx * tv[i]->a[c] / t.wc[i];
          dNy[c] += ly * tv[i]->a[c] / t.wc[i];
        }
      }
      const float ux = (dNx[0] - val[0] * dDx) / denom * cam.tex_size, uy = (dNy[0] - val[0] * dDy) / denom * cam.tex_size;
      const float vx = (dNx[1] - val[1] * dDx) / denom * cam.tex_size, vy = (dNy[1] - val[1] * dDy) / denom * cam.tex_size;
      const float rho = fmaxf(sqrtf(ux * ux + vx * vx), sqrtf(uy * uy + vy * vy));
      r_uv[0] = val[0];
      r_uv[1] = val[1];
      r_uv[2] = fminf(fmaxf(log2f(fmaxf(rho, 1e-30f)), 0.f), log2f(cam.tex_size));
      const float nn = rsqrtf(fmaxf(val[2] * val[2] + val[3] * val[3] + val[4] * val[4], 1e-30f));
      const float pn = rsqrtf(fmaxf(val[5] * val[5] + val[6] * val[6] + val[7] * val[7], 1e-30f));
      r_ang = fmaxf(-(val[2] * val[5] + val[3] * val[6] + val[4] * val[7]) * nn * pn, 0.f);
      // LinearizeDepth(gl_FragCoord.z) of depth.frag is the eye depth = 1 / (interpolated 1 / w_clip); taken from the
      // perspective interpolation rather than by inverting the float32 window depth (whose error grows with depth^2)
      r_dep = 1.f / denom;
    }
    uv_out[o] = r_uv[0]; uv_out[o + 1] = r_uv[1]; uv_out[o + 2] = r_uv[2];
    ang_out[o] = ang_out[o + 1] = ang_out[o + 2] = r_ang;
    dep_out[o] = dep_out[o + 1] = dep_out[o + 2] = r_dep;
  }
}

int launch_raster_view(const float* verts, int nv, const int* faces, int nf, const float* corner_uv,
                       const float* corner_normal, const float* view3x4, const float* proj6, int w, int h, float near,
                       float far, float tex_size, int flip, float* eye_scratch, unsigned long long* zbuf, float* uv_out,
                       float* angle_out, float* depth_out, cudaStream_t st) {
  SMB_REQUIRE(verts && faces && corner_uv && corner_normal && view3x4 && proj6 && eye_scratch && zbuf && uv_out &&
                  angle_out && depth_out, "raster_view: null argument");
  SMB_REQUIRE(nv > 0 && nf > 0 && w > 0 && h > 0 && far > near && near > 0.f && tex_size >= 1.f,
              "raster_view: need a non-empty mesh, a non-empty image and 0 < near < far");
  SMB_REQUIRE((long long)nf < (1LL << 30), "raster_view: at most 2^30 faces");
  RasterCam cam;
  for (int i = 0; i < 12; ++i) cam.V[i] = view3x4[i];
  cam.p00 = proj6[0]; cam.p02 = proj6[1]; cam.p11 = proj6[2]; cam.p12 = proj6[3]; cam.p22 = proj6[4]; cam.p23 = proj6[5];
  cam.w = w; cam.h = h; cam.flip = flip;
  cam.near = near; cam.far = far; cam.tex_size = tex_size;
  SMB_CUDA_CHECK(cudaMemsetAsync(zbuf, 0xff, (size_t)w * h * sizeof(unsigned long long), st));
  SMB_LAUNCH(raster_vertex_kernel, (unsigned)std::min(ceil_div(nv, 256), 148 * 8), 256, 0, st, verts, nv, cam,
             reinterpret_cast<float4*>(eye_scratch));
  SMB_LAUNCH(raster_zbuffer_kernel, (unsigned)std::min<long long>(ceil_div64((long long)nf * 32, 256), 148 * 16), 256, 0,
             st, reinterpret_cast<const float4*>(eye_scratch), faces, nf, cam, zbuf);
  SMB_LAUNCH(raster_resolve_kernel, (unsigned)std::min<long long>(ceil_div64((long long)w * h, 256), 148 * 8), 256, 0, st,
             reinterpret_cast<const float4*>(eye_scratch), faces, corner_uv, corner_normal, cam, zbuf, uv_out, angle_out,
             depth_out);
  return SMB_OK;
}

}  // namespace smb
