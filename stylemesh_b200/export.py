"""Texture export and headless preview on the device.

    texture_rgb8(texture)                  what the reference writes at every epoch end (model/model.py:378-385 ->
                                           texture.py:110-131): get_image() -> post() -> ToPILImage, as an RGB uint8
                                           (H, W, 3) device tensor, byte for byte what the reference hands to the JPEG
                                           encoder.  Only the H*W*3 bytes cross PCIe, not the fp32 texture.
    MipPreview(texture).render(uv[, bias]) the reference's post-run styled views (model/optimize.py:181-208 starts the
                                           OpenGL renderer with GL_LINEAR_MIPMAP_LINEAR on the exported texture) from the
                                           view's own (u, v, mip LOD) map, without a GL context.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _abi
from . import engine as _eng


def post_rgb8(bgr_chw: torch.Tensor) -> torch.Tensor:
    """post() + ToPILImage quantisation of a pre()-space (3,H,W) fp32 CUDA image -> (H,W,3) uint8 RGB (device)."""
    lib = _abi.load()
    x = _eng._require_cuda_f32(bgr_chw, "image")
    if x.dim() != 3 or x.shape[0] != 3:
        raise ValueError("post_rgb8 takes a (3, H, W) image")
    out = torch.empty((x.shape[1], x.shape[2], 3), device=x.device, dtype=torch.uint8)
    _abi.check(lib.smb_texture_post_rgb8(_abi.ptr(x), x.shape[1], x.shape[2], _abi.ptr(out), _abi.current_stream()),
               "smb_texture_post_rgb8")
    return out


def texture_rgb8(texture) -> torch.Tensor:
    """NeuralTexture / HierarchicalNeuralTexture -> the epoch-end texture image as (H,W,3) uint8 RGB on the device."""
    with torch.no_grad():
        img = texture.get_image()
        if not isinstance(img, torch.Tensor):
            raise TypeError("texture.get_image() must return a tensor")
        img = img.detach()
        if img.shape[0] != 3:                       # to_image(): first three channels, zero padded (texture.py:9-19)
            pad = torch.zeros(3 - img.shape[0], *img.shape[1:], device=img.device, dtype=img.dtype)
            img = torch.cat((img[:3], pad), dim=0)
        return post_rgb8(img[:3].contiguous().float())


def save_rgb8(rgb8: torch.Tensor, path: str) -> None:
    from PIL import Image
    Image.fromarray(rgb8.cpu().numpy()).save(path)


class MipPreview:
    """Mip chain of the composited texture (2x2 box filter per level, down to 1x1) + trilinear view lookup."""

    def __init__(self, texture, max_levels: int = 16):
        lib = _abi.load()
        with torch.no_grad():
            base = texture.get_image().detach()[:3].contiguous().float()
        _eng._require_cuda_f32(base, "texture image")
        self.levels: List[torch.Tensor] = [base]
        while len(self.levels) < max_levels and (self.levels[-1].shape[1] > 1 or self.levels[-1].shape[2] > 1):
            src = self.levels[-1]
            dst = torch.empty((3, max(src.shape[1] // 2, 1), max(src.shape[2] // 2, 1)), device=src.device,
                              dtype=torch.float32)
            _abi.check(lib.smb_mip_downsample2x(_abi.ptr(src), src.shape[1], src.shape[2], _abi.ptr(dst),
                                                _abi.current_stream()), "smb_mip_downsample2x")
            self.levels.append(dst)

    def render(self, uv: torch.Tensor, lod_bias: float = 0.0) -> torch.Tensor:
        """uv: (H, W, 2 or 3) fp32 in [0,1] as the renderer wrote it ([u, v, mip LOD]; (0,0) = no geometry) ->
        (H, W, 3) uint8 RGB on the device.  lod_bias shifts the stored LOD (log2 of the ratio between this texture's
        size and the size of the texture the UV maps were rendered with)."""
        lib = _abi.load()
        uv = _eng._require_cuda_f32(uv, "uv map")
        if uv.dim() != 3 or uv.shape[2] not in (2, 3):
            raise ValueError("uv must be (H, W, 2) or (H, W, 3)")
        H, W, ch = uv.shape
        out = torch.empty((H, W, 3), device=uv.device, dtype=torch.uint8)
        n = len(self.levels)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in self.levels])
        _abi.check(lib.smb_mip_preview(ptrs, _abi.int_array([t.shape[2] for t in self.levels]),
                                       _abi.int_array([t.shape[1] for t in self.levels]), n, _abi.ptr(uv), ch, H, W,
                                       float(lod_bias), _abi.ptr(out), _abi.current_stream()), "smb_mip_preview")
        return out


def grid_to_uv(grid: torch.Tensor) -> torch.Tensor:
    """grid_sample grid (H,W,2) in [-1,1] (the 13-tuple's uv entries) -> renderer uv (H,W,2) in [0,1]; the invalid
    marker (-1,-1) maps back to (0,0)."""
    return ((grid + 1.0) * 0.5).contiguous()
