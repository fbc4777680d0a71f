"""Learnable texture modules with the reference's API (model/texture/texture.py:9-135), backed by the sm_100a
sample / scatter kernels.

    NeuralTexture(W, H, C, random_init)              .data Parameter (C,H,W), .forward(grid (B,h,w,2)) -> (B,C,h,w)
    HierarchicalNeuralTexture(W, H, C, num_layers)   .layers ModuleList, forward = sum over layers, .regularizer()

Module-level `forward` is autograd-compatible (dense per-layer gradients, like grid_sampler_2d_backward).  The
training pipeline (stylemesh_b200/model/model.py) does not go through autograd: it scatters straight into a
persistent flat gradient buffer that the fused Adam kernel consumes.
"""
from __future__ import annotations

from os.path import join
from typing import List, Sequence

import torch
import torch.nn as nn

from ... import engine as _eng
from .utils import from_grid_range


def _save_texture_image(module, path, normalize_transform):
    """texture.py:59-65 / :123-127.  With the reference's post() transform on a CUDA texture the whole chain
    get_image -> post() -> ToPILImage runs on the device (stylemesh_b200.export.texture_rgb8: same bytes, and only
    H*W*3 bytes cross PCIe); any other transform takes the reference's host path."""
    img = module.get_image()
    if getattr(normalize_transform, "is_reference_post", False) and img.is_cuda:
        from ... import export as _export
        _export.save_rgb8(_export.texture_rgb8(module), path)
        return
    to_image(img, normalize_transform=normalize_transform).save(path)


def to_image(texture, startIndex=0, padChannels=True, normalize_transform=from_grid_range):
    """texture.py:9-19 — first three channels (zero padded) -> PIL image."""
    from torchvision.transforms import ToPILImage
    t = texture.detach().cpu()[startIndex:startIndex + 3].clone()
    if padChannels and t.shape[0] != 3:
        pad = torch.zeros(3 - t.shape[0], t.shape[1], t.shape[2], dtype=t.dtype)
        t = torch.cat((t, pad), dim=0)
    return ToPILImage()(normalize_transform(t))


class _UVSample(torch.autograd.Function):
    """out[b] = sum_l bilinear(clamp(layer_l), grid[b]); backward = UV scatter-add into dense per-layer grads."""

    @staticmethod
    def forward(ctx, grid, *layers):
        g = grid.detach()
        outs = [_eng.uv_sample_fwd([l.detach() for l in layers], g[b]) for b in range(g.shape[0])]
        ctx.save_for_backward(g)
        ctx.layer_shapes = [tuple(l.shape) for l in layers]
        return torch.stack(outs, dim=0)

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        grads = [torch.zeros(s, device=g.device, dtype=torch.float32) for s in ctx.layer_shapes]
        go = grad_out.contiguous()
        for b in range(g.shape[0]):
            _eng.uv_scatter_bwd(grads, g[b], go[b])
        return (None, *grads)


def sample_layers(layers: Sequence[torch.Tensor], grid: torch.Tensor) -> torch.Tensor:
    return _UVSample.apply(grid, *layers)


class NeuralTexture(nn.Module):
    def __init__(self, W, H, C, random_init=False):
        super().__init__()
        self.W, self.H, self.C = W, H, C
        init = torch.rand(C, H, W) if random_init else torch.zeros(C, H, W)      # texture.py:29-32
        self.data = nn.Parameter(init, requires_grad=True)

    @staticmethod
    def from_tensor(data: torch.Tensor):
        C, H, W = data.shape
        tex = NeuralTexture(W, H, C)
        tex.data = nn.Parameter(data, requires_grad=True)
        return tex

    def normalize(self):
        """texture.py:41-44 — the stored Parameter is clamped in place."""
        with torch.no_grad():
            self.data.clamp_(_eng.CLAMP_LO, _eng.CLAMP_HI)

    def forward(self, x):
        self.normalize()
        return sample_layers([self.data], x)

    def get_image(self):
        return self.data

    def save_image(self, dir, prefix="", normalize_transform=from_grid_range):
        _save_texture_image(self, join(dir, f"{prefix}texture.jpg"), normalize_transform)

    def save_layers(self, dir, prefix="", normalize_transform=from_grid_range):
        self.save_image(dir, prefix, normalize_transform)

    def save_texture(self, dir, prefix=""):
        torch.save(self.get_image().detach().cpu(), join(dir, f"{prefix}texture.pt"))


class HierarchicalNeuralTexture(nn.Module):
    def __init__(self, W, H, C, num_layers=4, random_init=False):
        super().__init__()
        self.W, self.H, self.C = W, H, C
        # Laplacian-style pyramid: layer i is (W // 2^i, H // 2^i)   (texture.py:80-81)
        self.layers = nn.ModuleList([NeuralTexture(W // 2 ** i, H // 2 ** i, C, random_init)
                                     for i in range(num_layers)])

    @staticmethod
    def from_tensor(data: list):
        C, H, W = data[0].shape
        for i, d in enumerate(data):
            if tuple(d.shape) != (C, H // 2 ** i, W // 2 ** i):
                raise AssertionError(f"layer {i} has shape {tuple(d.shape)}, expected {(C, H // 2 ** i, W // 2 ** i)}")
        tex = HierarchicalNeuralTexture(W, H, C, num_layers=len(data))
        tex.layers = nn.ModuleList([NeuralTexture.from_tensor(d) for d in data])
        return tex

    def layer_params(self) -> List[torch.Tensor]:
        return [l.data for l in self.layers]

    def forward(self, x):
        for l in self.layers:
            l.normalize()
        return sample_layers(self.layer_params(), x)        # one fused kernel over all layers (texture.py:96-100)

    def regularizer(self, weights):
        """texture.py:102-108 — sum_i w_i * mean(layer_i^2) (autograd-visible torch expression; the training
        pipeline uses the fused kernels instead)."""
        reg = 0.0
        for i, l in enumerate(self.layers):
            reg = reg + torch.mean(torch.pow(l.data, 2.0)) * weights[i]
        return reg

    def get_image(self):
        """texture.py:110-121 — dense identity-grid sample of the layer sum."""
        dev = self.layers[0].data.device
        w_range = torch.arange(0, self.W, dtype=torch.float, device=dev) / (self.W - 1.0) * 2.0 - 1.0
        h_range = torch.arange(0, self.H, dtype=torch.float, device=dev) / (self.H - 1.0) * 2.0 - 1.0
        v, u = torch.meshgrid(h_range, w_range, indexing="ij")
        uv_id = torch.stack([u, v], 2).unsqueeze(0).contiguous()
        with torch.no_grad():
            return self.forward(uv_id)[0, 0:3, :, :]

    def save_image(self, dir, prefix="", normalize_transform=from_grid_range):
        _save_texture_image(self, join(dir, f"{prefix}texture.jpg"), normalize_transform)

    def save_layers(self, dir, prefix="", normalize_transform=from_grid_range):
        for i, l in enumerate(self.layers):
            l.save_image(dir, prefix + f"_layer{i}_", normalize_transform)

    def save_texture(self, dir, prefix=""):
        for i, l in enumerate(self.layers):
            l.save_texture(dir, f"{prefix}layer-{i}-")
