"""VGG-19 style/content loss with the reference's module API (model/losses/content_and_style_losses.py), executed
by the sm_100a engine (tcgen05 implicit-GEMM convs, masked tcgen05 Gram, fused MSE, hand-written backward).

Public names kept: VGG, GramMatrix, image_pyramid, ContentAndStyleLoss (+ its class attributes, which
model/optimize.py reads for argparse defaults, reference optimize.py:274-285).

Differences that are deliberate and documented in DESIGN.md:
  * features are never compacted by boolean indexing (cs:136-143); the per-layer masks weight the Gram / MSE
    kernels instead — same numbers, no gather/scatter traffic;
  * only the 13 convs up to conv5_1 are evaluated (the reference also runs conv5_2..conv5_4 + pool5, unused);
  * gradients come from the engine's own backward (data-gradient convs, Gram backward, pool/ReLU masks), exposed
    to autograd through one torch.autograd.Function.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _abi
from ... import engine as _eng
from ...view_cache import ViewLRU

_ALL_CONVS = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
              ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256),
              ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512),
              ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512)]
_POOL_BEFORE = {2, 4, 8, 12}      # conv indices (of the 13) preceded by a 2x2 max-pool


def layer_hw(conv: int, H: int, W: int):
    """spatial size of relu(conv_i) for an H x W input (MaxPool2d(2,2) floor mode, cs:27-32)."""
    h, w = H, W
    for i in range(conv + 1):
        if i in _POOL_BEFORE:
            h, w = h // 2, w // 2
    return h, w


class VGG(nn.Module):
    """Frozen VGG-19 (Gatys weights layout, cs:7-45).  The nn.Conv2d children only hold the state_dict
    (`conv1_1.weight` ...) so that `vgg_conv.pth` loads unchanged; forward() runs on the B200 engine."""

    def __init__(self, pool="max", model_path=None, freeze=True):
        super().__init__()
        if pool != "max":
            raise NotImplementedError("the B200 engine implements MaxPool2d(2,2) only (every reference script uses it)")
        for name, cin, cout in _ALL_CONVS:
            setattr(self, name, nn.Conv2d(cin, cout, kernel_size=3, padding=1))
        if model_path:
            self.load_state_dict(torch.load(model_path, map_location="cpu"))
        if freeze:
            for p in self.parameters():
                p.requires_grad = False
        self._engine: Optional[_eng.VGGEngine] = None

    def engine(self) -> _eng.VGGEngine:
        if self._engine is None:
            self._engine = _eng.VGGEngine({k: v for k, v in self.state_dict().items()})
        return self._engine

    def reset_engine(self):
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    def forward(self, x, out_keys):
        """x (B,3,H,W) CUDA fp32 -> {key: (B,C,h,w)} for keys in r11..r51 (no autograd: weights are frozen and the
        training gradient flows through ContentAndStyleLoss)."""
        eng = self.engine()
        per_image = [eng.features(x[b], out_keys) for b in range(x.shape[0])]
        return {k: torch.cat([d[k] for d in per_image], dim=0) for k in out_keys}


class GramMatrix(nn.Module):
    """cs:74-80 — G = F F^T / (h*w) for F (b,c,h,w), on the tensor-core Gram kernel."""

    def forward(self, input):
        b, c, h, w = input.shape
        impl = _eng._impl_from_env("SMB_GRAM_IMPL", _abi.IMPL_TC)
        flat = input.reshape(b, c, h, w)
        return torch.stack([_eng.unit_gram(impl, flat[i], None, 1.0 / (h * w)) for i in range(b)], dim=0)


def image_pyramid(img, levels, reverse=False, minimum_size=256):
    """cs:83-133 — halving pyramid with a floor of `minimum_size` on the short side; one-off host-side prep."""
    h, w = img.shape[2:]
    entries, floor_img, floor_at = [], None, len(levels)
    for i, level in enumerate(levels):
        if level == 0:
            entries.append(img)
            continue
        hd, wd = int(h / 2 ** level), int(w / 2 ** level)
        if hd >= minimum_size and wd >= minimum_size:
            entries.append(F.interpolate(img, (hd, wd), mode="bilinear"))
            continue
        if floor_img is None:
            if w > h:
                size = (minimum_size, int(w * minimum_size / h))
            else:
                size = (int(h * minimum_size / w), minimum_size)
            floor_img = F.interpolate(img, size, mode="bilinear")
            floor_at = i
        entries.append(floor_img)
    if reverse:
        head = entries[:floor_at + 1][::-1]
        entries = head + [img] * (len(entries) - len(head))
    return entries


# ---------------------------------------------------------------------------------------------------------------
# per-view loss plan: masks, counts and pyramid factors (cs:146-217), built with a single host sync
# ---------------------------------------------------------------------------------------------------------------
class LossPlan:
    """Everything about a view's masks that the loss needs, as flat device row-masks plus host scalars."""

    def __init__(self):
        self.levels: List[dict] = []        # per level: {"size": (H,W), "layers": {name: {...}}}


def build_loss_plan(level_sizes, pyramid_masks, angle_degrees, angle_threshold, layer_names, need_angle_split):
    plan = LossPlan()
    stats = []
    for (H, W), mask in zip(level_sizes, pyramid_masks):
        entry = {"size": (H, W), "layers": {}}
        raw4 = mask.reshape(1, 1, H, W).float()
        m4 = (raw4 > 0).float()                            # {0,1}: the engine's row masks select pixels (cs:137 `mask > 0`)
        if need_angle_split:
            passed = F.interpolate(angle_degrees, (H, W), mode="bilinear") < angle_threshold      # cs:161
            m_pass4, m_fail4 = m4 * passed, m4 * (~passed)
        for name in layer_names:
            conv = _eng.layer_index(name)
            h, w = layer_hw(conv, H, W)
            m = F.interpolate(m4, (h, w), mode="nearest").reshape(-1).contiguous()                 # cs:172
            rec = {"conv": conv, "hw": (h, w), "mask": m}
            stats.append(m.sum())
            # cs:181: the level factor is the mean of the RAW nearest-resized mask values (differs from the pixel count
            # only for soft masks, which the public ContentAndStyleLoss.forward accepts)
            stats.append(F.interpolate(raw4, (h, w), mode="nearest").sum())
            if need_angle_split:
                rec["mask_pass"] = F.interpolate(m_pass4, (h, w), mode="nearest").reshape(-1).contiguous()
                rec["mask_fail"] = F.interpolate(m_fail4, (h, w), mode="nearest").reshape(-1).contiguous()
                stats.append(rec["mask_pass"].sum())
                stats.append(rec["mask_fail"].sum())
            entry["layers"][name] = rec
        plan.levels.append(entry)
    vals = torch.stack(stats).tolist() if stats else []       # the one host sync of the plan
    k = 0
    for entry in plan.levels:
        for name in layer_names:
            rec = entry["layers"][name]
            rec["n"] = vals[k]; k += 1
            raw_sum = vals[k]; k += 1
            if need_angle_split:
                rec["n_pass"] = vals[k]; k += 1
                rec["n_fail"] = vals[k]; k += 1
            rec["f_raw"] = raw_sum / float(rec["hw"][0] * rec["hw"][1])                             # cs:181
    for name in layer_names:                                                                        # cs:199-204
        total = sum(e["layers"][name]["f_raw"] for e in plan.levels)
        for e in plan.levels:
            e["layers"][name]["f"] = (e["layers"][name]["f_raw"] / total) if total > 0 else float("nan")
    return plan


def loss_plan_from_counts(level_sizes, per_level, counts_per_level, layer_names, need_angle_split):
    """LossPlan from the mask-pyramid kernel's outputs (engine.view_level_plan): per_level[i]["layers"][k] holds the
    row masks of layer_names[k], counts_per_level[i] = [alive, n_0, n_pass_0, n_fail_0, n_1, ...] (host ints).
    Same fields and the same factor normalisation as build_loss_plan (cs:181,199-204); the level masks are binary, so
    the raw mean of cs:181 equals n / (h * w)."""
    plan = LossPlan()
    for (H, W), lv, cnt in zip(level_sizes, per_level, counts_per_level):
        entry = {"size": (H, W), "layers": {}}
        for k, name in enumerate(layer_names):
            conv = _eng.layer_index(name)
            h, w = layer_hw(conv, H, W)
            rec = {"conv": conv, "hw": (h, w), "mask": lv["layers"][k]["mask"], "n": float(cnt[1 + 3 * k])}
            if need_angle_split:
                rec["mask_pass"] = lv["layers"][k]["mask_pass"]
                rec["mask_fail"] = lv["layers"][k]["mask_fail"]
                rec["n_pass"], rec["n_fail"] = float(cnt[2 + 3 * k]), float(cnt[3 + 3 * k])
            rec["f_raw"] = rec["n"] / float(h * w)
            entry["layers"][name] = rec
        plan.levels.append(entry)
    for name in layer_names:
        total = sum(e["layers"][name]["f_raw"] for e in plan.levels)
        for e in plan.levels:
            e["layers"][name]["f"] = (e["layers"][name]["f_raw"] / total) if total > 0 else float("nan")
    return plan


def _inv(n: float) -> float:
    return 1.0 / n if n > 0 else 0.0


class ContentAndStyleLoss(nn.Module):
    # same defaults as cs:221-238 (read by the CLI for argparse defaults)
    style_layers = ["r11", "r21", "r31", "r41", "r51"]
    content_layers = ["r42"]
    style_weights = [1e3 / n ** 2 for n in [64, 128, 256, 512, 512]]
    content_weights = [1 for _ in range(len(content_layers))]
    style_pyramid_modes = ["single", "multi"]
    gram_modes = ["current", "average"]

    def __init__(self, model_path, style_layers=style_layers, content_layers=content_layers,
                 style_weights=style_weights, content_weights=content_weights, angle_threshold=60,
                 style_pyramid_mode="single", gram_mode="current"):
        super().__init__()
        if not model_path:
            raise ValueError("No model_path provided")
        self.vgg = VGG(model_path=model_path)
        self.style_layers = list(style_layers)
        self.content_layers = list(content_layers)
        self.layers = self.style_layers + self.content_layers
        for name in self.layers:
            _eng.layer_index(name)                      # ValueError for layers the engine does not compute
        self.style_weights = list(style_weights)
        self.content_weights = list(content_weights)
        if style_pyramid_mode not in self.style_pyramid_modes:
            raise ValueError(f"Unsupported style_pyramid_mode: {style_pyramid_mode}")
        if gram_mode not in self.gram_modes:
            raise ValueError(f"Unsupported gram_mode: {gram_mode}")
        self.style_pyramid_mode = style_pyramid_mode
        self.gram_mode = gram_mode
        self.gram_cache = {k: [] for k in self.style_layers}
        self.style_targets = None
        self.angle_threshold = angle_threshold
        self.cache_content_targets = False             # opt-in: reuse VGG(target) per view index
        self._content_cache = ViewLRU(16 << 30)        # VGG(target) features per view, least recently used evicted

    # -- style targets (one-off, cs:273-286) ----------------------------------------------------------------
    def set_style_image(self, style_image, num_levels=5):
        levels = list(range(num_levels))
        pyramid = image_pyramid(style_image, levels, reverse=True)
        print("Use style image pyramid of shapes:")
        for p in pyramid:
            print(p.shape)
        eng = self.vgg.engine()
        convs = [_eng.layer_index(n) for n in self.style_layers]
        per_entry = {}
        for p in pyramid:                          # entries repeat (the floor image / the original): encode once
            key = (id(p), tuple(p.shape))
            if key in per_entry:
                continue
            img = p[0].to(eng.device, torch.float32).contiguous()
            slot = eng.begin(img.shape[1], img.shape[2])
            eng.forward(slot, img, max(convs), keep=convs)      # style targets: Grams of these layers only
            per_entry[key] = []
            for c in convs:
                _, h, w = eng.feature_shape(slot, c)
                per_entry[key].append(eng.gram(slot, c, None, 1.0 / (h * w)).unsqueeze(0))
        self.style_targets = [{l: per_entry[(id(pyramid[k]), tuple(pyramid[k].shape))][i] for k, l in enumerate(levels)}
                              for i in range(len(self.style_layers))]
        eng.release_slots()                        # style images can be 2048 px: give the memory back

    # -- content targets ------------------------------------------------------------------------------------
    def content_targets(self, target_content, level_sizes, cache_key=None):
        """VGG(target)[content layers] at the target's own resolution (cs:294), then bilinear-resized to every
        level's layer size (cs:176) and laid out channels-last for the content kernel."""
        if self.cache_content_targets and cache_key is not None:
            hit = self._content_cache.get(cache_key)
            if hit is not None:
                return hit
        eng = self.vgg.engine()
        out = {}
        if self.content_layers:
            img = target_content[0].contiguous()
            convs = [_eng.layer_index(n) for n in self.content_layers]
            slot = eng.begin(img.shape[1], img.shape[2])
            eng.forward(slot, img, max(convs), keep=convs)      # never back-propagated: inference-only pass
            for name, c in zip(self.content_layers, convs):
                C_, hc, wc = eng.feature_shape(slot, c)
                nhwc = None
                per_level = []
                for (H, W) in level_sizes:
                    h, w = layer_hw(c, H, W)
                    if (h, w) == (hc, wc):
                        if nhwc is None:
                            nhwc = eng.feature_nhwc(slot, c)
                        per_level.append(nhwc)
                    else:
                        t = F.interpolate(eng.feature(slot, c).unsqueeze(0), (h, w), mode="bilinear")
                        per_level.append(t[0].permute(1, 2, 0).reshape(h * w, C_).contiguous())
                out[name] = per_level
        if self.cache_content_targets and cache_key is not None:
            self._content_cache.put(cache_key, out)
        return out

    # -- the fused evaluation: losses + d(loss)/d(pred) in one pass ------------------------------------------
    def fused_loss_and_grads(self, preds: Sequence[torch.Tensor], plan: LossPlan, content_tgts: dict,
                             style_scale: float, content_scale: float, loss_accum: torch.Tensor,
                             want_grads: bool = True, gram_history: Optional[dict] = None,
                             record_history: Optional[dict] = None):
        """preds[i]: (3,H_i,W_i).  Adds style_scale*style_loss to loss_accum[0] and content_scale*content_loss to
        loss_accum[1]; returns [d(sum)/d(pred_i)] (already scaled) or None.
        gram_mode='average' (cs:319-323): every (level, layer) term averages its Gram with the <= 9 most recent cached
        ones - including those of EARLIER LEVELS OF THIS CALL, the cache is shared across levels - and pushes its own.
        record_history={} receives {(level, layer): (prev_sum, avg_len)} as used; gram_history=<that dict> replays a
        call with exactly those histories and leaves the cache alone (the autograd backward)."""
        if self.style_targets is None:
            raise RuntimeError("set_style_image() must be called before the loss is evaluated")
        eng = self.vgg.engine()
        multi = self.style_pyramid_mode == "multi"
        deepest = max(_eng.layer_index(n) for n in self.layers)
        grads = []
        acc_style, acc_content = loss_accum[0:1], loss_accum[1:2]
        for li, (pred, entry) in enumerate(zip(preds, plan.levels)):
            H, W = entry["size"]
            slot = eng.begin(H, W)
            eng.forward(slot, pred, deepest)
            for idx, name in enumerate(self.style_layers):
                rec = entry["layers"][name]
                coef = style_scale * self.style_weights[idx] * rec["f"]
                tgt = self.style_targets[idx]
                prev_sum, avg_len, gram_out = None, 1.0, None
                if self.gram_mode == "average":                                                  # cs:319-323
                    if gram_history is not None:
                        prev_sum, avg_len = gram_history[(li, name)]
                    else:
                        prev = self.gram_cache[name][:9]
                        avg_len = float(len(prev) + 1)
                        prev_sum = torch.stack(prev).sum(0).contiguous() if prev else None
                        if record_history is not None:
                            record_history[(li, name)] = (prev_sum, avg_len)
                    gram_out = torch.empty_like(tgt[0][0])
                if multi:                                                                        # cs:305-338
                    eng.style_term(slot, rec["conv"], rec["mask_pass"], _inv(rec["n_pass"]), tgt[2][0], coef,
                                   tgt[0][0] if idx > 2 else None, coef, acc_style, prev_sum, avg_len, gram_out)
                    if rec["n_fail"] > 0:
                        eng.style_term(slot, rec["conv"], rec["mask_fail"], _inv(rec["n_fail"]), tgt[2][0], coef,
                                       None, 0.0, acc_style)
                else:
                    eng.style_term(slot, rec["conv"], rec["mask"], _inv(rec["n"]), tgt[0][0], coef, None, 0.0,
                                   acc_style, prev_sum, avg_len, gram_out)
                if self.gram_mode == "average" and gram_history is None:
                    self.gram_cache[name] = [gram_out] + self.gram_cache[name][:9]
            for idx, name in enumerate(self.content_layers):                                     # cs:343-348
                rec = entry["layers"][name]
                if rec["n"] <= 0:
                    if rec["f"] != rec["f"]:            # cs:199-204: a zero factor sum makes the reference's term NaN
                        acc_content += float("nan")
                    continue
                C_ = _eng.CONV_COUT[rec["conv"]]
                coef_loss = content_scale * self.content_weights[idx] * rec["f"] / (C_ * rec["n"])
                eng.content_term(slot, rec["conv"], content_tgts[name][li], rec["mask"], coef_loss, 2.0 * coef_loss,
                                 acc_content)
            if want_grads:
                grads.append(eng.backward(slot, H, W))
        return grads if want_grads else None

    # -- reference-compatible forward (autograd) ---------------------------------------------------------------
    def forward(self, pred, target_content, pyramid_masks, angle_unnormalized=None):
        """pred: list of (1,3,H_i,W_i); pyramid_masks: list of (1,1,H_i,W_i); returns (style (1,), content (1,),
        pyramid-info dict) like cs:288-350.  Batch size 1 (the reference's masked_features breaks for B>1)."""
        sizes = [tuple(p.shape[2:]) for p in pred]
        plan = build_loss_plan(sizes, pyramid_masks, angle_unnormalized, self.angle_threshold, self.layers,
                               self.style_pyramid_mode == "multi")
        tgts = self.content_targets(target_content, sizes)
        style, content = _LossFunction.apply(self, plan, tgts, *pred)
        info = {"f": [{k: e["layers"][k]["f"] for k in self.layers} for e in plan.levels],
                "m": [{k: e["layers"][k]["mask"] for k in self.layers} for e in plan.levels],
                "size": len(plan.levels)}
        return style, content, info


class _LossFunction(torch.autograd.Function):
    """Values in forward; backward re-evaluates the step with the incoming scalar gradients folded into the term
    coefficients (the VGG forwards are re-run: another evaluation may have reused the per-resolution slots in
    between) and returns the engine's own data gradients.  The Gram histories of gram_mode='average' are the ones the
    forward used (stored on ctx), not a re-slice of the live cache."""

    @staticmethod
    def forward(ctx, module: ContentAndStyleLoss, plan, tgts, *preds):
        acc = torch.zeros(2, device=preds[0].device, dtype=torch.float32)
        imgs = [p.detach()[0].contiguous() for p in preds]
        history = {}
        module.fused_loss_and_grads(imgs, plan, tgts, 1.0, 1.0, acc, want_grads=False, record_history=history)
        ctx.module, ctx.plan, ctx.tgts, ctx.history = module, plan, tgts, history
        ctx.save_for_backward(*imgs)
        return acc[0:1].clone(), acc[1:2].clone()

    @staticmethod
    def backward(ctx, g_style, g_content):
        gs, gc = torch.stack([g_style.reshape(-1)[0], g_content.reshape(-1)[0]]).tolist()
        scratch = torch.zeros(2, device=g_style.device, dtype=torch.float32)
        grads = ctx.module.fused_loss_and_grads(list(ctx.saved_tensors), ctx.plan, ctx.tgts, gs, gc, scratch,
                                                want_grads=True, gram_history=ctx.history)
        return (None, None, None, *[g.unsqueeze(0) for g in grads])
