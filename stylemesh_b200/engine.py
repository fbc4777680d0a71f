"""Thin Python handles over the C-ABI (include/stylemesh_b200.h): texture ops and the VGG/loss engine.

torch is used for device memory, streams and tiny shape glue only; every arithmetic step of the hot path is a
kernel in libstylemesh_b200.so.  Nothing here falls back to torch ops or to the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _abi

CLAMP_LO, CLAMP_HI = -123.6800, 151.0610     # model/texture/texture.py:43

# relu(conv) names of model/losses/content_and_style_losses.py:49-66 -> index into the 13-conv table
LAYER_INDEX = {"r11": 0, "r12": 1, "r21": 2, "r22": 3, "r31": 4, "r32": 5, "r33": 6, "r34": 7,
               "r41": 8, "r42": 9, "r43": 10, "r44": 11, "r51": 12}
CONV_NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv3_4",
              "conv4_1", "conv4_2", "conv4_3", "conv4_4", "conv5_1"]
CONV_COUT = [64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512]


def layer_index(name: str) -> int:
    if name not in LAYER_INDEX:
        raise ValueError(f"Unsupported VGG layer '{name}': the B200 engine computes relu outputs r11..r51 "
                         f"({sorted(LAYER_INDEX)})")
    return LAYER_INDEX[name]


def require_cuda_device(device) -> None:
    if torch.device(device).type != "cuda":
        raise _abi.StyleMeshB200Error(
            f"tensors live on '{device}': move the module to a CUDA device (model.cuda()); stylemesh_b200 has no CPU path")


def _require_cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _abi.StyleMeshB200Error(f"{what} must be a CUDA tensor: stylemesh_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ---------------------------------------------------------------------------------------------------------------
# texture ops
# ---------------------------------------------------------------------------------------------------------------
def uv_sample_fwd(layers: Sequence[torch.Tensor], grid: torch.Tensor, out: Optional[torch.Tensor] = None,
                  clamp=(CLAMP_LO, CLAMP_HI)) -> torch.Tensor:
    """layers: list of (C,H_l,W_l); grid (h,w,2) -> out (C,h,w).  (texture.py:46-54, 96-100)"""
    lib = _abi.load()
    layers = [_require_cuda_f32(t, "texture layer") for t in layers]
    grid = _require_cuda_f32(grid, "uv grid")
    h, w = grid.shape[-3], grid.shape[-2]
    c = layers[0].shape[0]
    if out is None:
        out = torch.empty((c, h, w), device=grid.device, dtype=torch.float32)
    rc = lib.smb_uv_sample_fwd(_abi.ptr_array(layers), _abi.int_array([t.shape[2] for t in layers]),
                               _abi.int_array([t.shape[1] for t in layers]), len(layers), c, _abi.ptr(grid), h, w,
                               clamp[0], clamp[1], _abi.ptr(out), _abi.current_stream())
    _abi.check(rc, "smb_uv_sample_fwd")
    return out


def uv_texel_index(grid: torch.Tensor, tex_w: int, tex_h: int):
    lib = _abi.load()
    grid = _require_cuda_f32(grid, "uv grid")
    n = grid.numel() // 2
    xy0 = torch.empty((n, 2), device=grid.device, dtype=torch.int32)
    w4 = torch.empty((n, 4), device=grid.device, dtype=torch.float32)
    _abi.check(lib.smb_uv_texel_index(_abi.ptr(grid), n, tex_w, tex_h, _abi.ptr(xy0), _abi.ptr(w4),
                                      _abi.current_stream()), "smb_uv_texel_index")
    return xy0, w4


def uv_scatter_bwd(grad_layers: Sequence[torch.Tensor], grid: torch.Tensor, grad_out: torch.Tensor,
                   hook0: Optional[torch.Tensor] = None, hook1: Optional[torch.Tensor] = None) -> None:
    """grad_layers[l] (C,H_l,W_l) += scatter(grad_out (C,h,w) * hook0 * hook1)   (accumulates)."""
    lib = _abi.load()
    for t in grad_layers:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("gradient layers must be contiguous CUDA float32 tensors (they are written in place)")
    grid = _require_cuda_f32(grid, "uv grid")
    grad_out = _require_cuda_f32(grad_out, "grad_out")
    h, w = grid.shape[-3], grid.shape[-2]
    c = grad_layers[0].shape[0]
    hook0 = None if hook0 is None else _require_cuda_f32(hook0, "hook0")
    hook1 = None if hook1 is None else _require_cuda_f32(hook1, "hook1")
    rc = lib.smb_uv_scatter_bwd(_abi.ptr_array(grad_layers), _abi.int_array([t.shape[2] for t in grad_layers]),
                                _abi.int_array([t.shape[1] for t in grad_layers]), len(grad_layers), c,
                                _abi.ptr(grid), h, w, _abi.ptr(grad_out), _abi.ptr(hook0), _abi.ptr(hook1),
                                _abi.current_stream())
    _abi.check(rc, "smb_uv_scatter_bwd")


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, reg_coef=0.0, grad_scale=1.0,
              clamp=(CLAMP_LO, CLAMP_HI)) -> None:
    lib = _abi.load()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("adam_step operates in place on contiguous CUDA float32 tensors")
    rc = lib.smb_adam_step(_abi.ptr(param), _abi.ptr(grad), _abi.ptr(exp_avg), _abi.ptr(exp_avg_sq), param.numel(),
                           lr, beta1, beta2, eps, int(step), clamp[0], clamp[1], reg_coef, grad_scale,
                           _abi.current_stream())
    _abi.check(rc, "smb_adam_step")


def _segment_arrays(seg_begin, seg_coef):
    import ctypes as C
    n = len(seg_begin)
    if n != len(seg_coef) or n < 1:
        raise ValueError("segment offsets and coefficients must have the same non-zero length")
    return (C.c_int64 * n)(*[int(b) for b in seg_begin]), (C.c_float * n)(*[float(c) for c in seg_coef]), n


def adam_step_segments(param, grad, exp_avg, exp_avg_sq, seg_begin, seg_reg_coef, lr, beta1, beta2, eps, step,
                       grad_scale=1.0, clamp=(CLAMP_LO, CLAMP_HI)) -> None:
    """adam_step over a flat buffer of consecutive layers (segment l starts at seg_begin[l]) in one launch."""
    lib = _abi.load()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("adam_step_segments operates in place on contiguous CUDA float32 tensors")
    begin, coef, n = _segment_arrays(seg_begin, seg_reg_coef)
    rc = lib.smb_adam_step_segments(_abi.ptr(param), _abi.ptr(grad), _abi.ptr(exp_avg), _abi.ptr(exp_avg_sq),
                                    param.numel(), begin, coef, n, lr, beta1, beta2, eps, int(step), clamp[0], clamp[1],
                                    grad_scale, _abi.current_stream())
    _abi.check(rc, "smb_adam_step_segments")


def dist_adam_step(rank: int, world: int, grad_ptrs, param_ptrs, flag_ptrs, exp_avg, exp_avg_sq, numel: int, seg_begin,
                   seg_reg_coef, lr, beta1, beta2, eps, step, epoch: int, clamp=(CLAMP_LO, CLAMP_HI)) -> None:
    """Gradient reduce-scatter + Adam on this rank's slice + parameter all-gather over peer memory (one kernel) and
    the local gradient reset; *_ptrs are lists of `world` device addresses (every rank's buffer, own included)."""
    import ctypes as C
    lib = _abi.load()
    if not (len(grad_ptrs) == len(param_ptrs) == len(flag_ptrs) == world):
        raise ValueError("need one gradient / parameter / flag pointer per rank")
    for t in (exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == numel):
            raise ValueError("dist_adam_step: the moment buffers must be contiguous CUDA float32 of the flat size")
    arr = lambda ps: (C.c_void_p * world)(*[int(p) for p in ps])
    begin, coef, n = _segment_arrays(seg_begin, seg_reg_coef)
    rc = lib.smb_dist_adam_step(int(rank), int(world), arr(grad_ptrs), arr(param_ptrs), arr(flag_ptrs),
                                _abi.ptr(exp_avg), _abi.ptr(exp_avg_sq), int(numel), begin, coef, n, lr, beta1, beta2,
                                eps, int(step), clamp[0], clamp[1], int(epoch) & 0xFFFFFFFF, _abi.current_stream())
    _abi.check(rc, "smb_dist_adam_step")


def texreg_value_segments(param: torch.Tensor, seg_begin, seg_coef, out_accum: torch.Tensor,
                          clamp=(CLAMP_LO, CLAMP_HI)) -> None:
    """out_accum += sum_l seg_coef[l] * sum(clamp(segment l)^2) in one launch."""
    lib = _abi.load()
    begin, coef, n = _segment_arrays(seg_begin, seg_coef)
    rc = lib.smb_texreg_value_segments(_abi.ptr(param), param.numel(), begin, coef, n, clamp[0], clamp[1],
                                       _abi.ptr(out_accum), _abi.current_stream())
    _abi.check(rc, "smb_texreg_value_segments")


def launch_count() -> int:
    """kernel launches issued by libstylemesh_b200.so in this process so far."""
    return int(_abi.load().smb_launch_count())


def texreg_value(param: torch.Tensor, coef: float, out_accum: torch.Tensor, clamp=(CLAMP_LO, CLAMP_HI)) -> None:
    lib = _abi.load()
    rc = lib.smb_texreg_value(_abi.ptr(param), param.numel(), coef, clamp[0], clamp[1], _abi.ptr(out_accum),
                              _abi.current_stream())
    _abi.check(rc, "smb_texreg_value")


# ---------------------------------------------------------------------------------------------------------------
# view preparation (data/abstract_dataset.py:270-344 on the device; used by stylemesh_b200.data.ViewStore)
# ---------------------------------------------------------------------------------------------------------------
def _require_cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")
    return t


def view_uv_to_grid(uv_hw3: torch.Tensor, want_mask: bool = False, depth_at_uv: Optional[torch.Tensor] = None):
    """(H,W,3) renderer UV map -> grid (H,W,2) in [-1,1] and, optionally, the validity mask (H,W) bool."""
    lib = _abi.load()
    uv = _require_cuda(uv_hw3, torch.float32, "uv")
    H, W, c = uv.shape
    if c != 3:
        raise ValueError("the renderer's UV maps have 3 channels (u, v, mip LOD)")
    grid = torch.empty((H, W, 2), device=uv.device, dtype=torch.float32)
    valid = torch.empty((H, W), device=uv.device, dtype=torch.uint8) if want_mask else None
    if depth_at_uv is not None:
        _require_cuda(depth_at_uv, torch.float64, "depth_at_uv")
    _abi.check(lib.smb_view_uv_to_grid(_abi.ptr(uv), H, W, _abi.ptr(grid), _abi.ptr(valid), _abi.ptr(depth_at_uv),
                                       _abi.current_stream()), "smb_view_uv_to_grid")
    return grid, (valid.bool() if want_mask else None)


def view_gather2d(src: torch.Tensor, ytab: torch.Tensor, xtab: torch.Tensor) -> torch.Tensor:
    """dst[y][x] = src[ytab[y]][xtab[x]] (2-D, 1- or 4-byte elements; tables: CUDA int32)."""
    lib = _abi.load()
    if not (src.is_cuda and src.dim() == 2 and src.is_contiguous() and src.element_size() in (1, 4)):
        raise ValueError("view_gather2d takes a contiguous 2-D CUDA tensor of 1- or 4-byte elements")
    _require_cuda(ytab, torch.int32, "ytab")
    _require_cuda(xtab, torch.int32, "xtab")
    dst = torch.empty((ytab.numel(), xtab.numel()), device=src.device, dtype=src.dtype)
    _abi.check(lib.smb_view_gather2d(_abi.ptr(src), src.element_size(), src.shape[0], src.shape[1], _abi.ptr(ytab),
                                     _abi.ptr(xtab), dst.shape[0], dst.shape[1], _abi.ptr(dst), _abi.current_stream()),
               "smb_view_gather2d")
    return dst


_DEPTH_TYPES = {torch.float64: 0, torch.float32: 1, torch.uint16: 2, torch.int16: 2}


def view_resize_linear(src: torch.Tensor, out_hw, tables=None, divisor: float = 1.0) -> torch.Tensor:
    """cv2 INTER_LINEAR of a 2-D depth map into float64; tables = (yofs, yalpha, xofs, xalpha) CUDA tensors built
    by stylemesh_b200.data.resample (not needed when the size does not change)."""
    lib = _abi.load()
    if not (src.is_cuda and src.dim() == 2 and src.is_contiguous() and src.dtype in _DEPTH_TYPES):
        raise ValueError("view_resize_linear takes a contiguous 2-D CUDA tensor (float64, float32 or uint16)")
    Hd, Wd = int(out_hw[0]), int(out_hw[1])
    same = (Hd, Wd) == tuple(src.shape)
    if not same and tables is None:
        raise ValueError("resampling tables are required when the size changes")
    yo, ya, xo, xa = tables if tables is not None else (None, None, None, None)
    dst = torch.empty((Hd, Wd), device=src.device, dtype=torch.float64)
    _abi.check(lib.smb_view_resize_linear(_abi.ptr(src), _DEPTH_TYPES[src.dtype], float(divisor), src.shape[0],
                                          src.shape[1], _abi.ptr(yo), _abi.ptr(ya), _abi.ptr(xo), _abi.ptr(xa), Hd, Wd,
                                          _abi.ptr(dst), _abi.current_stream()), "smb_view_resize_linear")
    return dst


def view_depth_levels(depth: torch.Tensor, levels, min_depth: float, depth_is_f32: bool = False):
    """-> (depth_level f32, depth f32, rounded i64, other i64, weight f32), each shaped like `depth` (float64)."""
    import ctypes as C
    lib = _abi.load()
    d = _require_cuda(depth, torch.float64, "depth")
    lv = (C.c_double * len(levels))(*[float(x) for x in levels])
    outs = [torch.empty(d.shape, device=d.device, dtype=t)
            for t in (torch.float32, torch.float32, torch.int64, torch.int64, torch.float32)]
    _abi.check(lib.smb_view_depth_levels(_abi.ptr(d), d.numel(), lv, len(levels), float(min_depth), int(depth_is_f32),
                                         _abi.ptr(outs[0]), _abi.ptr(outs[1]), _abi.ptr(outs[2]), _abi.ptr(outs[3]),
                                         _abi.ptr(outs[4]), _abi.current_stream()), "smb_view_depth_levels")
    return tuple(outs)


def view_rgb_pre(rgb_hwc_u8: torch.Tensor) -> torch.Tensor:
    lib = _abi.load()
    x = _require_cuda(rgb_hwc_u8, torch.uint8, "rgb")
    H, W, c = x.shape
    if c != 3:
        raise ValueError("rgb must be (H, W, 3) uint8")
    out = torch.empty((3, H, W), device=x.device, dtype=torch.float32)
    _abi.check(lib.smb_view_rgb_pre(_abi.ptr(x), H, W, _abi.ptr(out), _abi.current_stream()), "smb_view_rgb_pre")
    return out


def view_angle_degrees(cos_angle: torch.Tensor) -> torch.Tensor:
    lib = _abi.load()
    x = _require_cuda(cos_angle, torch.float32, "cos_angle")
    out = torch.empty_like(x)
    _abi.check(lib.smb_view_angle_degrees(_abi.ptr(x), x.numel(), _abi.ptr(out), _abi.current_stream()),
               "smb_view_angle_degrees")
    return out


def view_erode3x3(x: torch.Tensor) -> torch.Tensor:
    """model/model.py:204-208 on a (..., H, W) float32 map with at most one non-singleton leading dimension."""
    lib = _abi.load()
    x = _require_cuda(x, torch.float32, "x")
    H, W = x.shape[-2], x.shape[-1]
    if x.numel() != H * W:
        raise ValueError("view_erode3x3 takes one (H, W) map (leading dimensions of size 1 are allowed)")
    out = torch.empty_like(x)
    _abi.check(lib.smb_view_erode3x3(_abi.ptr(x), H, W, _abi.ptr(out), _abi.current_stream()), "smb_view_erode3x3")
    return out


def view_level_masks(mask: torch.Tensor, rounded: torch.Tensor, other: torch.Tensor, interp_w: torch.Tensor,
                     num_levels: int):
    """model/model.py:210-239 for ALL pyramid levels at the rgb resolution in one launch.
    mask (H,W) bool/uint8, rounded / other (H,W) int64, interp_w (H,W) f32 ->
    (level_mask (L,H,W) f32 = erode(((rounded == l) | (other == l)) & mask),
     level_weight (L,H,W) f32 = erode((rounded == l) & mask) * w + erode((other == l) & mask) * (1 - w))."""
    lib = _abi.load()
    m = mask.to(torch.uint8) if mask.dtype != torch.uint8 else mask
    m = _require_cuda(m.contiguous(), torch.uint8, "mask")
    r = _require_cuda(rounded.contiguous(), torch.int64, "rounded_depth_level")
    o = _require_cuda(other.contiguous(), torch.int64, "other_depth_level")
    w = _require_cuda(interp_w.contiguous(), torch.float32, "depth_level_interpolation_weight")
    H, W = m.shape[-2], m.shape[-1]
    if not (m.numel() == r.numel() == o.numel() == w.numel() == H * W):
        raise ValueError("view_level_masks takes (H, W) maps of one view")
    lm = torch.empty((num_levels, H, W), device=m.device, dtype=torch.float32)
    lw = torch.empty((num_levels, H, W), device=m.device, dtype=torch.float32)
    _abi.check(lib.smb_view_level_masks(_abi.ptr(m), _abi.ptr(r), _abi.ptr(o), _abi.ptr(w), H, W, int(num_levels),
                                        _abi.ptr(lm), _abi.ptr(lw), _abi.current_stream()), "smb_view_level_masks")
    return lm, lw


def view_level_plan(src_mask: torch.Tensor, src_weight: Optional[torch.Tensor], angle_guidance: Optional[torch.Tensor],
                    angle_degrees: Optional[torch.Tensor], threshold: float, level_hw, layer_hw, counts: torch.Tensor):
    """One pyramid level in one launch (model.py:199,219,238,253-254 + cs:161,172-185): returns
    {"hook0": (H*W,) | None, "hook1": (H*W,) | None, "layers": [{"mask", "mask_pass"?, "mask_fail"?}, ...]} and fills
    `counts` (uint32/int32 device tensor of 1 + 3*len(layer_hw)): selected level pixels, then (n, n_pass, n_fail) per
    layer.  src_* are (Hr, Wr) maps at the rgb resolution; angle_degrees=None skips the pass / fail split."""
    lib = _abi.load()
    sm = _require_cuda(src_mask, torch.float32, "src_mask")
    Hr, Wr = sm.shape[-2], sm.shape[-1]
    H, W = int(level_hw[0]), int(level_hw[1])
    for t, name in ((src_weight, "src_weight"), (angle_guidance, "angle_guidance"), (angle_degrees, "angle_degrees")):
        if t is not None:
            _require_cuda(t, torch.float32, name)
            if t.numel() != Hr * Wr:
                raise ValueError(f"{name} must have the rgb resolution {Hr}x{Wr}")
    split = angle_degrees is not None
    nl = len(layer_hw)
    if counts.numel() < 1 + 3 * nl or counts.element_size() != 4 or not counts.is_cuda:
        raise ValueError("counts must be a CUDA 32-bit integer tensor of 1 + 3 * len(layer_hw) entries")
    hook0 = torch.empty(H * W, device=sm.device, dtype=torch.float32) if angle_guidance is not None else None
    hook1 = torch.empty(H * W, device=sm.device, dtype=torch.float32) if src_weight is not None else None
    per = 3 if split else 1
    sizes = [int(h) * int(w) for h, w in layer_hw]
    buf = torch.empty(per * sum(sizes), device=sm.device, dtype=torch.float32)
    _abi.check(lib.smb_view_level_plan(_abi.ptr(sm), _abi.ptr(src_weight), _abi.ptr(angle_guidance),
                                       _abi.ptr(angle_degrees), float(threshold), Hr, Wr, H, W, _abi.ptr(hook0),
                                       _abi.ptr(hook1), nl, _abi.int_array([h for h, _ in layer_hw]),
                                       _abi.int_array([w for _, w in layer_hw]), _abi.ptr(buf), int(split),
                                       _abi.ptr(counts), _abi.current_stream()), "smb_view_level_plan")
    layers, off = [], 0
    for n in sizes:
        rec = {"mask": buf[off:off + n]}
        off += n
        if split:
            rec["mask_pass"] = buf[off:off + n]
            rec["mask_fail"] = buf[off + n:off + 2 * n]
            off += 2 * n
        layers.append(rec)
    return {"hook0": hook0, "hook1": hook1, "layers": layers}


# ---------------------------------------------------------------------------------------------------------------
# VGG / loss engine
# ---------------------------------------------------------------------------------------------------------------
def _impl_from_env(name: str, default: int) -> int:
    v = os.environ.get(name, "").strip().lower()
    if v in ("", "default"):
        return default
    if v in ("tc", "tcgen05", "1"):
        return _abi.IMPL_TC
    if v in ("ph", "tc_ph", "tc5", "5"):
        return _abi.IMPL_TC_PH
    if v in ("simt", "0"):
        return _abi.IMPL_SIMT
    raise ValueError(f"{name}={v!r}: expected 'ph' (pair + halo tcgen05 convs), 'tc' (stream-K tcgen05) or 'simt'")


class VGGEngine:
    """Owns an smb_ctx: frozen VGG-19 conv1_1..conv5_1 in kernel layout + per-resolution working sets."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], conv_impl: Optional[int] = None,
                 gram_impl: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _abi.StyleMeshB200Error("VGGEngine needs a CUDA device: stylemesh_b200 has no CPU path")
        self._lib = _abi.load()
        torch.cuda.current_device()          # make sure the primary context exists and is current
        torch.zeros(1, device="cuda")
        self._ctx = self._lib.smb_ctx_create()
        if not self._ctx:
            raise _abi.StyleMeshB200Error(f"smb_ctx_create failed: {_abi.last_error()}")
        self.conv_impl = _impl_from_env("SMB_CONV_IMPL", _abi.IMPL_TC_PH) if conv_impl is None else conv_impl
        self.gram_impl = _impl_from_env("SMB_GRAM_IMPL", _abi.IMPL_TC) if gram_impl is None else gram_impl
        _abi.check(self._lib.smb_ctx_set_impl(self._ctx, self.conv_impl, self.gram_impl), "smb_ctx_set_impl")
        ws, bs = [], []
        for name in CONV_NAMES:
            ws.append(state_dict[name + ".weight"].detach().to("cpu", torch.float32).contiguous())
            bs.append(state_dict[name + ".bias"].detach().to("cpu", torch.float32).contiguous())
        _abi.check(self._lib.smb_ctx_load_vgg(self._ctx, _abi.ptr_array(ws), _abi.ptr_array(bs), len(ws)),
                   "smb_ctx_load_vgg")
        self.device = torch.device("cuda", torch.cuda.current_device())

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.smb_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- slots --------------------------------------------------------------------------------------------
    def begin(self, H: int, W: int) -> int:
        return _abi.check(self._lib.smb_level_begin(self._ctx, int(H), int(W)), "smb_level_begin")

    def release_slots(self) -> None:
        _abi.check(self._lib.smb_ctx_release_slots(self._ctx), "smb_ctx_release_slots")

    def device_bytes(self) -> int:
        return int(self._lib.smb_ctx_device_bytes(self._ctx))

    # -- per-kernel-class timing (bench.py roofline pass) ------------------------------------------------------
    TIMING_CLASSES = ["conv1_1_fwd", "igemm_conv_fwd", "igemm_conv_dgrad", "igemm_gram_bwd", "gram", "gram_mse",
                      "pool", "conv1_1_dgrad", "content_mse", "misc"]

    def set_timing(self, enabled: bool) -> None:
        _abi.check(self._lib.smb_ctx_set_timing(self._ctx, int(bool(enabled))), "smb_ctx_set_timing")

    def read_timing(self) -> Dict[str, dict]:
        n = len(self.TIMING_CLASSES)
        ms, fl, cnt = (C.c_float * n)(), (C.c_double * n)(), (C.c_int * n)()
        _abi.check(self._lib.smb_ctx_read_timing(self._ctx, ms, fl, cnt, n), "smb_ctx_read_timing")
        return {name: {"ms": float(ms[i]), "flops": float(fl[i]), "launches": int(cnt[i])}
                for i, name in enumerate(self.TIMING_CLASSES)}

    # -- forward / features ---------------------------------------------------------------------------------
    def forward(self, slot: int, image: torch.Tensor, last_conv: int, keep=None) -> None:
        """keep=None: training forward.  keep=[conv indices]: inference-only pass that guarantees only those layers'
        features (layers that merely feed a pool are pooled in the conv epilogue and never materialised)."""
        image = _require_cuda_f32(image, "image")
        if keep is None:
            _abi.check(self._lib.smb_level_forward(self._ctx, slot, _abi.ptr(image), int(last_conv),
                                                   _abi.current_stream()), "smb_level_forward")
            return
        mask = 0
        for c in keep:
            mask |= 1 << int(c)
        _abi.check(self._lib.smb_level_forward_features(self._ctx, slot, _abi.ptr(image), int(last_conv), mask,
                                                        _abi.current_stream()), "smb_level_forward_features")

    def feature_shape(self, slot: int, conv: int):
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        _abi.check(self._lib.smb_level_feature_shape(self._ctx, slot, conv, C.byref(c), C.byref(h), C.byref(w)),
                   "smb_level_feature_shape")
        return c.value, h.value, w.value

    def feature(self, slot: int, conv: int) -> torch.Tensor:
        c, h, w = self.feature_shape(slot, conv)
        out = torch.empty((c, h, w), device=self.device, dtype=torch.float32)
        _abi.check(self._lib.smb_level_get_feature(self._ctx, slot, conv, _abi.ptr(out), _abi.current_stream()),
                   "smb_level_get_feature")
        return out

    def feature_nhwc(self, slot: int, conv: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        c, h, w = self.feature_shape(slot, conv)
        if out is None:
            out = torch.empty((h * w, c), device=self.device, dtype=torch.float32)
        _abi.check(self._lib.smb_level_get_feature_nhwc(self._ctx, slot, conv, _abi.ptr(out),
                                                        _abi.current_stream()), "smb_level_get_feature_nhwc")
        return out

    def gram(self, slot: int, conv: int, rowmask: Optional[torch.Tensor], inv_n: float) -> torch.Tensor:
        c = CONV_COUT[conv]
        out = torch.empty((c, c), device=self.device, dtype=torch.float32)
        _abi.check(self._lib.smb_level_gram(self._ctx, slot, conv, _abi.ptr(rowmask), float(inv_n), _abi.ptr(out),
                                            _abi.current_stream()), "smb_level_gram")
        return out

    # -- loss terms / backward ------------------------------------------------------------------------------
    def style_term(self, slot: int, conv: int, rowmask: Optional[torch.Tensor], inv_n: float,
                   target0: torch.Tensor, coef0: float, target1: Optional[torch.Tensor], coef1: float,
                   loss_accum: torch.Tensor, prev_sum: Optional[torch.Tensor] = None, avg_len: float = 1.0,
                   gram_out: Optional[torch.Tensor] = None) -> None:
        _abi.check(self._lib.smb_level_style_term(self._ctx, slot, conv, _abi.ptr(rowmask), float(inv_n),
                                                  _abi.ptr(target0), float(coef0), _abi.ptr(target1), float(coef1),
                                                  _abi.ptr(prev_sum), float(avg_len), _abi.ptr(gram_out),
                                                  _abi.ptr(loss_accum), _abi.current_stream()),
                   "smb_level_style_term")

    def content_term(self, slot: int, conv: int, target_nhwc: torch.Tensor, rowmask: torch.Tensor,
                     coef_loss: float, coef_grad: float, loss_accum: torch.Tensor) -> None:
        _abi.check(self._lib.smb_level_content_term(self._ctx, slot, conv, _abi.ptr(target_nhwc), _abi.ptr(rowmask),
                                                    float(coef_loss), float(coef_grad), _abi.ptr(loss_accum),
                                                    _abi.current_stream()), "smb_level_content_term")

    def backward(self, slot: int, H: int, W: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((3, H, W), device=self.device, dtype=torch.float32)
        _abi.check(self._lib.smb_level_backward(self._ctx, slot, _abi.ptr(out), _abi.current_stream()),
                   "smb_level_backward")
        return out

    # -- whole-network convenience (VGG.forward of the reference) --------------------------------------------
    def features(self, image: torch.Tensor, keys: Sequence[str]) -> Dict[str, torch.Tensor]:
        """image (3,H,W) or (1,3,H,W) -> {key: (1,C,h,w)}"""
        img = image[0] if image.dim() == 4 else image
        idx = [layer_index(k) for k in keys]
        slot = self.begin(img.shape[1], img.shape[2])
        self.forward(slot, img, max(idx))
        return {k: self.feature(slot, i).unsqueeze(0) for k, i in zip(keys, idx)}


# ---------------------------------------------------------------------------------------------------------------
# unit-level wrappers (parity tests)
# ---------------------------------------------------------------------------------------------------------------
def unit_conv3x3(impl: int, x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool,
                 transpose_flip: bool = False) -> torch.Tensor:
    lib = _abi.load()
    x = _require_cuda_f32(x, "x")
    cout, cin = w.shape[0], w.shape[1]
    _, H, W = x.shape
    wc = w.detach().to("cpu", torch.float32).contiguous()
    bc = None if b is None else b.detach().to("cpu", torch.float32).contiguous()
    y = torch.empty((cin if transpose_flip else cout, H, W), device=x.device, dtype=torch.float32)
    _abi.check(lib.smb_unit_conv3x3(impl, _abi.ptr(x), cin, H, W, _abi.ptr(wc), _abi.ptr(bc), cout, int(relu),
                                    int(transpose_flip), _abi.ptr(y), _abi.current_stream()), "smb_unit_conv3x3")
    return y


def unit_conv3x3_fused(x: torch.Tensor, w: torch.Tensor, f: torch.Tensor, g: torch.Tensor,
                       rowmask: Optional[torch.Tensor]) -> torch.Tensor:
    """conv3x3(x, w) + rowmask * (g @ f) on the pair + halo kernel (the Gram backward folded into a data-gradient conv)."""
    lib = _abi.load()
    x = _require_cuda_f32(x, "x")
    f = _require_cuda_f32(f, "f")
    cout, cin = w.shape[0], w.shape[1]
    _, H, W = x.shape
    wc = w.detach().to("cpu", torch.float32).contiguous()
    gc = g.detach().to("cpu", torch.float32).contiguous()
    y = torch.empty((cout, H, W), device=x.device, dtype=torch.float32)
    _abi.check(lib.smb_unit_conv3x3_fused(_abi.ptr(x), cin, H, W, _abi.ptr(wc), cout, _abi.ptr(f), _abi.ptr(gc),
                                          _abi.ptr(rowmask), _abi.ptr(y), _abi.current_stream()),
               "smb_unit_conv3x3_fused")
    return y


def unit_maxpool(x: torch.Tensor) -> torch.Tensor:
    lib = _abi.load()
    x = _require_cuda_f32(x, "x")
    c, H, W = x.shape
    y = torch.empty((c, H // 2, W // 2), device=x.device, dtype=torch.float32)
    _abi.check(lib.smb_unit_maxpool(_abi.ptr(x), c, H, W, _abi.ptr(y), _abi.current_stream()), "smb_unit_maxpool")
    return y


def unit_maxpool_bwd(g: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    lib = _abi.load()
    g = _require_cuda_f32(g, "g")
    y = _require_cuda_f32(y, "y")
    c, H, W = y.shape
    dx = torch.empty_like(y)
    _abi.check(lib.smb_unit_maxpool_bwd(_abi.ptr(g), _abi.ptr(y), c, H, W, _abi.ptr(dx), _abi.current_stream()),
               "smb_unit_maxpool_bwd")
    return dx


def unit_gram(impl: int, f: torch.Tensor, rowmask: Optional[torch.Tensor], inv_n: float) -> torch.Tensor:
    lib = _abi.load()
    f = _require_cuda_f32(f, "f")
    c, H, W = f.shape
    G = torch.empty((c, c), device=f.device, dtype=torch.float32)
    _abi.check(lib.smb_unit_gram(impl, _abi.ptr(f), c, H, W, _abi.ptr(rowmask), float(inv_n), _abi.ptr(G),
                                 _abi.current_stream()), "smb_unit_gram")
    return G
