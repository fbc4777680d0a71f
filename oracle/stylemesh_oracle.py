"""CPU oracle of the StyleMesh per-view texture-optimisation step.  TEST INFRASTRUCTURE ONLY.

This file restates, function by function, what the reference (lukasHoel/stylemesh, /root/reference) computes on the
hot path, using plain torch fp32 CPU ops — the same third-party ops the reference itself calls (torch is the
reference's arithmetic library: requirements.txt:98 pins torch==1.9.1; 2.11 is installed here and has identical
semantics for the arguments used).  It exists to CHECK the CUDA path and to be TIMED as the
CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import it;
the product package stylemesh_b200 never does and has no CPU fallback.

Parity pinning: the reference ships no tests and no golden vectors ("parity unpinned" by its own tests).
The oracle is therefore pinned against outputs of the REAL reference modules imported from /root/reference in the
build container — tests/golden/make_golden.py generates tests/golden/*.pt, tests/test_oracle_vs_golden.py checks
this file against them.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------------------
# constants
# ----------------------------------------------------------------------------------------------------------------
CLAMP_LO, CLAMP_HI = -123.6800, 151.0610                      # model/texture/texture.py:43
IMAGENET_MEAN_BGR = (0.40760392, 0.45795686, 0.48501961)      # model/losses/rgb_transform.py:7

# model/losses/content_and_style_losses.py:11-26 (all 16 convs) and :27-32 (pools after blocks 1..5)
VGG_CONV_SPECS = [
    ("conv1_1", 3, 64), ("conv1_2", 64, 64),
    ("conv2_1", 64, 128), ("conv2_2", 128, 128),
    ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256),
    ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512),
    ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512),
]
_BLOCK_SIZES = (2, 2, 4, 4, 4)
DEFAULT_STYLE_LAYERS = ["r11", "r21", "r31", "r41", "r51"]     # cs:222
DEFAULT_CONTENT_LAYERS = ["r42"]                               # cs:223
DEFAULT_STYLE_WEIGHTS = [1e3 / n ** 2 for n in [64, 128, 256, 512, 512]]   # cs:226


# ----------------------------------------------------------------------------------------------------------------
# colour / grid transforms
# ----------------------------------------------------------------------------------------------------------------
def pre_transform(rgb01: torch.Tensor) -> torch.Tensor:
    """rgb_transform.py:5-11 — RGB->BGR, subtract ImageNet mean, x255. (3,H,W) in [0,1] -> VGG space."""
    bgr = rgb01[[2, 1, 0]].clone()
    mean = torch.tensor(IMAGENET_MEAN_BGR, dtype=bgr.dtype).view(3, 1, 1)
    return (bgr - mean) * 255.0


def post_transform(x: torch.Tensor) -> torch.Tensor:
    """rgb_transform.py:14-21 — inverse of pre_transform plus clamp to [0,1] (never in place)."""
    y = x.detach().cpu().clone() * (1.0 / 255.0)
    mean = torch.tensor(IMAGENET_MEAN_BGR, dtype=y.dtype).view(3, 1, 1)
    y = y + mean
    return y[[2, 1, 0]].clamp(0, 1)


def to_grid(uv_chw: torch.Tensor) -> torch.Tensor:
    """model/texture/utils.py:6-8,21-23,56-60 — uv in [0,1] (C>=2,H,W) -> grid_sample grid (H,W,2) in [-1,1]."""
    g = uv_chw * 2.0 - 1
    return g[:2].permute(1, 2, 0)


# ----------------------------------------------------------------------------------------------------------------
# texture
# ----------------------------------------------------------------------------------------------------------------
def texture_normalize_(layers: Sequence[torch.Tensor]) -> None:
    """texture.py:41-44 — in-place clamp of every layer Parameter."""
    with torch.no_grad():
        for t in layers:
            t.copy_(torch.clamp(t, CLAMP_LO, CLAMP_HI))


def texture_sample(layers: Sequence[torch.Tensor], grid: torch.Tensor) -> torch.Tensor:
    """texture.py:46-54 and :96-100 — sum over layers of grid_sample(bilinear, border, align_corners=True).
    layers: list of (C,H_l,W_l); grid (B,h,w,2).  Returns (B,C,h,w)."""
    texture_normalize_(layers)
    b = grid.shape[0]
    ys = [F.grid_sample(t.repeat(b, 1, 1, 1) if t.dim() == 4 else t.unsqueeze(0).repeat(b, 1, 1, 1), grid,
                        mode="bilinear", padding_mode="border", align_corners=True) for t in layers]
    if len(ys) == 1:
        return ys[0]
    return torch.sum(torch.stack(ys), dim=0)


def texture_regularizer(layers: Sequence[torch.Tensor], weights: Sequence[float]) -> torch.Tensor:
    """texture.py:102-108 — sum_i mean(layer_i^2) * w_i."""
    reg = 0.0
    for i, t in enumerate(layers):
        reg = reg + torch.mean(torch.pow(t, 2.0)) * weights[i]
    return reg


def uv_texel_indices_np(grid: np.ndarray, W: int, H: int):
    """ATen/native/GridSampler.h:27-36 (unnormalize, align_corners=True), :58-60 (clip), :143-171 — the exact fp32
    index arithmetic, one IEEE operation at a time in numpy float32.  grid (...,2) float32.
    Returns x0,y0 (int32) and weights (...,4) float32 in the order nw,ne,sw,se."""
    g = grid.astype(np.float32)
    one, half = np.float32(1.0), np.float32(0.5)

    def unnorm(c, size):
        t = (c + one).astype(np.float32)
        t = (t * half).astype(np.float32)          # "/ 2" is exact
        t = (t * np.float32(size - 1)).astype(np.float32)
        return np.minimum(np.float32(size - 1), np.maximum(t, np.float32(0.0))).astype(np.float32)

    ix, iy = unnorm(g[..., 0], W), unnorm(g[..., 1], H)
    fx, fy = np.floor(ix).astype(np.float32), np.floor(iy).astype(np.float32)
    x1, y1 = (fx + one).astype(np.float32), (fy + one).astype(np.float32)
    ax, bx = (x1 - ix).astype(np.float32), (ix - fx).astype(np.float32)
    ay, by = (y1 - iy).astype(np.float32), (iy - fy).astype(np.float32)
    w = np.stack([(ax * ay), (bx * ay), (ax * by), (bx * by)], axis=-1).astype(np.float32)
    return fx.astype(np.int32), fy.astype(np.int32), w


# ----------------------------------------------------------------------------------------------------------------
# VGG-19
# ----------------------------------------------------------------------------------------------------------------
def vgg_forward(params: Dict[str, torch.Tensor], x: torch.Tensor, out_keys: Sequence[str],
                as_written: bool = True) -> Dict[str, torch.Tensor]:
    """cs:47-70 — relu(conv3x3 pad 1) x16 with 2x2 max-pools after blocks 1..5.
    as_written=True runs all 16 convs + pool5 like the reference does on every call; False stops after the deepest
    requested key (identical values for the requested keys)."""
    out: Dict[str, torch.Tensor] = {}
    deepest = None
    if not as_written:
        order = []
        for b, n in enumerate(_BLOCK_SIZES, start=1):
            order += [f"r{b}{k}" for k in range(1, n + 1)] + [f"p{b}"]
        deepest = max(order.index(k) for k in out_keys)
        pos = 0
    h = x
    for b, n in enumerate(_BLOCK_SIZES, start=1):
        for k in range(1, n + 1):
            name = f"conv{b}_{k}"
            h = F.relu(F.conv2d(h, params[name + ".weight"], params[name + ".bias"], padding=1))
            out[f"r{b}{k}"] = h
            if deepest is not None:
                if pos == deepest:
                    return {kk: out[kk] for kk in out_keys}
                pos += 1
        h = F.max_pool2d(h, kernel_size=2, stride=2)
        out[f"p{b}"] = h
        if deepest is not None:
            if pos == deepest:
                return {kk: out[kk] for kk in out_keys}
            pos += 1
    return {k: out[k] for k in out_keys}


def gram_matrix(f: torch.Tensor) -> torch.Tensor:
    """cs:74-80 — bmm(F, F^T) / (h*w) for F (b,c,h,w)."""
    b, c, h, w = f.shape
    fl = f.reshape(b, c, h * w)
    return torch.bmm(fl, fl.transpose(1, 2)) / (h * w)


def image_pyramid(img: torch.Tensor, levels: Sequence[int], reverse: bool = False,
                  minimum_size: int = 256) -> List[torch.Tensor]:
    """cs:83-133 — halving pyramid that never drops below `minimum_size` on the short side; with reverse=True the
    entries up to the first floor entry are reversed and the rest padded with the original image."""
    h, w = img.shape[2:]
    pyr: List[torch.Tensor] = []
    floor_entry = None
    floor_index = len(levels)
    for i, level in enumerate(levels):
        if level == 0:
            pyr.append(img)
            continue
        hd, wd = int(h / 2 ** level), int(w / 2 ** level)
        if hd < minimum_size or wd < minimum_size:
            if floor_entry is None:
                if w > h:
                    fh = minimum_size
                    fw = int(w * fh / h)
                else:
                    fw = minimum_size
                    fh = int(h * fw / w)
                floor_entry = F.interpolate(img, (fh, fw), mode="bilinear")
                floor_index = i
            pyr.append(floor_entry)
        else:
            pyr.append(F.interpolate(img, (hd, wd), mode="bilinear"))
    if reverse:
        head = pyr[:floor_index + 1][::-1]
        while len(head) < len(pyr):
            head.append(img)
        pyr = head
    return pyr


def masked_features(f: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """cs:136-143 — compact the valid pixels: (1,C,h,w) -> (1,C,Nvalid,1); all-invalid -> zeros (1,C,h*w,1)."""
    sel = f[:, :, mask.squeeze() > 0].unsqueeze(3)
    if sel.shape[2] == 0:
        return torch.zeros_like(f).reshape(f.shape[0], f.shape[1], -1).unsqueeze(3)
    return sel


@dataclass
class StyleContentOracle:
    """cs:220-350 — ContentAndStyleLoss restated."""
    vgg_params: Dict[str, torch.Tensor]
    style_layers: List[str] = field(default_factory=lambda: list(DEFAULT_STYLE_LAYERS))
    content_layers: List[str] = field(default_factory=lambda: list(DEFAULT_CONTENT_LAYERS))
    style_weights: List[float] = field(default_factory=lambda: list(DEFAULT_STYLE_WEIGHTS))
    content_weights: List[float] = field(default_factory=lambda: [1.0])
    angle_threshold: float = 60.0
    style_pyramid_mode: str = "single"
    gram_mode: str = "current"
    as_written: bool = True
    style_targets: Optional[list] = None
    gram_cache: Dict[str, list] = field(default_factory=dict)

    def __post_init__(self):
        self.gram_cache = {k: [] for k in self.style_layers}

    @property
    def layers(self):
        return self.style_layers + self.content_layers

    def set_style_image(self, style_image: torch.Tensor, num_levels: int = 5) -> None:
        """cs:273-286 — style_targets[layer_index][level] = Gram(VGG(pyramid[level])[layer])."""
        levels = list(range(num_levels))
        pyr = image_pyramid(style_image, levels, reverse=True)
        with torch.no_grad():
            enc = [vgg_forward(self.vgg_params, p, self.style_layers, self.as_written) for p in pyr]
            self.style_targets = [{l: gram_matrix(enc[k][name]).detach() for k, l in enumerate(levels)}
                                  for name in self.style_layers]

    def loss(self, pred: List[torch.Tensor], target_content: torch.Tensor, pyramid_masks: List[torch.Tensor],
             angle_degrees: torch.Tensor):
        """cs:288-350 (forward) with calculate_pyramid cs:146-217 inlined.  Returns (style (1,), content (1,))."""
        enc = [vgg_forward(self.vgg_params, p, self.layers, self.as_written) for p in pred]          # cs:291
        with torch.no_grad():
            content_enc = vgg_forward(self.vgg_params, target_content, self.layers, self.as_written)  # cs:294
        n_levels = len(enc)
        factors: List[Dict[str, torch.Tensor]] = []
        feats, feats_pass, feats_fail, fail_masks, content_tgt = [], [], [], [], []
        for li, e in enumerate(enc):
            mask = pyramid_masks[li]
            passed = F.interpolate(angle_degrees, mask.shape[2:], mode="bilinear") < self.angle_threshold  # cs:161
            f_i, a_i, p_i, q_i, m_fail_i, c_i = {}, {}, {}, {}, {}, {}
            for k, o in e.items():
                with torch.no_grad():                                                               # cs:171-185
                    m = F.interpolate(mask, o.shape[2:], mode="nearest")
                    m_pass = F.interpolate(mask * passed, o.shape[2:], mode="nearest")
                    m_fail = F.interpolate(mask * (~passed), o.shape[2:], mode="nearest")
                    c_i[k] = masked_features(F.interpolate(content_enc[k], o.shape[2:], mode="bilinear"), m)
                    f_i[k] = torch.mean(m)
                    m_fail_i[k] = m_fail
                a_i[k] = masked_features(o, m)                                                      # cs:187-189
                p_i[k] = masked_features(o, m_pass)
                q_i[k] = masked_features(o, m_fail)
            factors.append(f_i)
            feats.append(a_i)
            feats_pass.append(p_i)
            feats_fail.append(q_i)
            fail_masks.append(m_fail_i)
            content_tgt.append(c_i)
        for k in self.layers:                                                                        # cs:199-204
            total = sum(factors[i][k] for i in range(n_levels))
            for i in range(n_levels):
                factors[i][k] = factors[i][k] / total

        style = torch.zeros(1)
        content = torch.zeros(1)
        mse = F.mse_loss
        for li in range(n_levels):                                                                   # cs:301
            for idx, name in enumerate(self.style_layers):                                           # cs:304
                if self.style_pyramid_mode == "single":
                    y = self.style_targets[idx][0]
                    y_hat = gram_matrix(feats[li][name])
                elif self.style_pyramid_mode == "multi":
                    y = self.style_targets[idx][2]
                    y_hat = gram_matrix(feats_pass[li][name])
                else:
                    raise ValueError(f"Unsupported style_pyramid_mode: {self.style_pyramid_mode}")
                if self.gram_mode == "average":                                                      # cs:319-323
                    cache = [g.detach() for g in self.gram_cache[name][:9]]
                    cache.insert(0, y_hat)
                    self.gram_cache[name] = cache
                    y_hat = torch.mean(torch.stack(cache), dim=0)
                f = factors[li][name]
                l = self.style_weights[idx] * f * mse(y, y_hat)                                      # cs:326
                if self.style_pyramid_mode == "multi":                                               # cs:328-338
                    g_fail = gram_matrix(feats_fail[li][name])
                    if torch.sum(fail_masks[li][name]) > 0:
                        l = l + self.style_weights[idx] * f * mse(y, g_fail)
                    if idx > 2:
                        l = l + self.style_weights[idx] * f * mse(self.style_targets[idx][0], y_hat)
                style = style + l
            for idx, name in enumerate(self.content_layers):                                         # cs:343-348
                f = factors[li][name]
                content = content + self.content_weights[idx] * f * mse(content_tgt[li][name], feats[li][name])
        return style, content


# ----------------------------------------------------------------------------------------------------------------
# step glue (model/model.py:178-270) and optimiser (model/model.py:387-401)
# ----------------------------------------------------------------------------------------------------------------
def erode(x: torch.Tensor, kernel_size: int = 3) -> torch.Tensor:
    """model.py:204-208 — keep x where the 3x3 box mean equals 1."""
    k = torch.ones(1, 1, kernel_size, kernel_size, dtype=x.dtype)
    e = torch.clamp(F.conv2d(x, k, padding=(1, 1)) / kernel_size ** 2, 0, 1)
    return x * (e == 1)


def level_masks_and_weights(batch, level_sizes, use_depth_scaling: bool):
    """model.py:210-254 — per-level loss masks and depth interpolation weights (None when depth scaling is off).
    level_sizes: list of (h,w) of the sampled predictions."""
    (_, _, _, _, _, rounded, other, interp_w, _, _, mask, _, _) = batch
    mask_f = mask.unsqueeze(1).float()
    masks, weights = [], []
    if use_depth_scaling:
        for i, size in enumerate(level_sizes):
            m = (((rounded == i) + (other == i)).float()) * mask_f                                   # :211-214
            m = F.interpolate(erode(m), size, mode="nearest")
            masks.append((m > 0).float())
            m1 = erode((rounded == i) * mask_f) * interp_w                                          # :225-233
            m2 = erode((other == i) * mask_f) * (1 - interp_w)
            weights.append(F.interpolate(m1 + m2, size, mode="nearest"))
    else:
        for size in level_sizes:                                                                     # :253
            masks.append((F.interpolate(torch.zeros_like(mask_f), size, mode="nearest") > 0).float())
        masks[-1] = (F.interpolate(mask_f, level_sizes[-1], mode="nearest") > 0).float()            # :254
        weights = None
    return masks, weights


@dataclass
class OracleConfig:
    """ctor arguments of TextureOptimizationStyleTransferPipeline that influence the step (model.py:25-60)."""
    use_angle_weight: bool = True
    use_depth_scaling: bool = True
    loss_weights: Dict[str, float] = field(default_factory=lambda: {"content": 0.0, "style": 0.0, "tex_reg": 0.0})
    tex_reg_weights: Optional[List[float]] = None
    hierarchical: bool = True
    learning_rate: float = 1e-3
    decay_gamma: float = 0.1
    decay_step_size: int = 30


class OraclePipeline:
    """The reference's LightningModule step + optimizer, functional style."""

    def __init__(self, layers: List[torch.Tensor], loss: StyleContentOracle, cfg: OracleConfig):
        self.layers = [t.detach().clone().requires_grad_(True) for t in layers]
        self.loss_fn = loss
        self.cfg = cfg
        if cfg.hierarchical and not cfg.tex_reg_weights:                                             # model.py:85-88
            n = len(layers)
            cfg.tex_reg_weights = [float(pow(2, n - i - 1)) for i in range(n)]
            cfg.tex_reg_weights[-1] = 0
        if cfg.hierarchical and len(layers) != len(cfg.tex_reg_weights):                             # model.py:90-92
            raise ValueError(f"Have {len(layers)} texture layers, but only {len(cfg.tex_reg_weights)} weights specified")
        self.opt = torch.optim.Adam([{"params": self.layers, "weight_decay": 0.0, "lr": cfg.learning_rate}],
                                    lr=cfg.learning_rate, weight_decay=0.0)                          # model.py:391-395
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, gamma=cfg.decay_gamma, step_size=cfg.decay_step_size)

    def forward_with_loss(self, batch):
        """model.py:178-270 — returns dict of weighted losses (content, style, tex_reg, total)."""
        (rgb, _, _, _, _, _, _, _, _, uvs, mask, angle_guidance, angle_degrees) = batch
        cfg = self.cfg
        pred = [texture_sample(self.layers, v) for v in uvs]                                         # model.py:157-159
        if cfg.use_angle_weight:                                                                     # :195-202
            for p in pred:
                p.register_hook(lambda g: g * F.interpolate(angle_guidance, g.shape[2:], mode="bilinear"))
        masks, weights = level_masks_and_weights(batch, [p.shape[2:] for p in pred], cfg.use_depth_scaling)
        if cfg.use_depth_scaling:                                                                    # :245-251
            by_h = {w.shape[2]: w for w in weights}
            for p in pred:
                p.register_hook(lambda g: g * by_h[g.shape[2]])
        keep = [i for i, m in enumerate(masks) if torch.sum(m) > 0]                                  # :256-257
        pred_k = [pred[i] for i in keep]
        masks_k = [masks[i] for i in keep]
        style, content = self.loss_fn.loss(pred_k, rgb, masks_k, angle_degrees)                      # :259
        lw = cfg.loss_weights
        losses = {"content": lw["content"] * content, "style": lw["style"] * style}                 # :261-262
        if lw.get("tex_reg", 0.0) > 0:                                                               # :264-267
            if cfg.hierarchical:
                losses["tex_reg"] = lw["tex_reg"] * texture_regularizer(self.layers, cfg.tex_reg_weights)
            else:
                losses["tex_reg"] = lw["tex_reg"] * torch.zeros(1)
        else:
            losses["tex_reg"] = torch.zeros_like(losses["content"])
        losses["total"] = sum(losses.values())                                                       # :270
        return losses

    def grads(self, batch):
        """loss terms + dense texture gradients for one view, no optimiser update (teacher-forced parity)."""
        for t in self.layers:
            t.grad = None
        losses = self.forward_with_loss(batch)
        losses["total"].backward()
        g = [t.grad.detach().clone() if t.grad is not None else torch.zeros_like(t) for t in self.layers]
        return {k: float(v.detach().reshape(-1)[0]) for k, v in losses.items()}, g

    def step(self, batch):
        """one Lightning automatic-optimisation step: zero_grad, backward, Adam.step."""
        self.opt.zero_grad()
        losses = self.forward_with_loss(batch)
        losses["total"].backward()
        self.opt.step()
        return {k: float(v.detach().reshape(-1)[0]) for k, v in losses.items()}

    def step_views(self, batches):
        """N-rank data-parallel oracle: mean of the per-view loss gradients, regulariser once, one
        Adam step — what N view-sharded ranks + all-reduce(mean) compute."""
        self.opt.zero_grad()
        n = len(batches)
        out = None
        reg_w = self.cfg.loss_weights.get("tex_reg", 0.0)
        for b in batches:
            losses = self.forward_with_loss(b)
            # the regulariser is replicated, not averaged: every rank adds the same term once
            data_term = losses["content"] + losses["style"]
            (data_term / n).backward()
            vals = {k: float(v.detach().reshape(-1)[0]) for k, v in losses.items()}
            out = vals if out is None else {k: out[k] + vals[k] for k in vals}
        if reg_w > 0 and self.cfg.hierarchical:
            (reg_w * texture_regularizer(self.layers, self.cfg.tex_reg_weights)).backward()
        self.opt.step()
        return {k: v / n for k, v in out.items()}

    def end_epoch(self):
        self.sched.step()                                                                            # StepLR per epoch
