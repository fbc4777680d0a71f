"""A CPU emulation of the C-ABI engine's SEMANTICS (include/stylemesh_b200.h) built from the oracle's torch ops.
TEST INFRASTRUCTURE: it lets the host-side logic of the product (loss plan, coefficients, fused-state plumbing,
optimizer, DDP scaling) be checked against the golden fixtures in the GPU-less container.  It is monkeypatched in
by tests only; the product has no way to select it."""
import torch
import torch.nn.functional as F

from oracle import stylemesh_oracle as orc

CLAMP = (orc.CLAMP_LO, orc.CLAMP_HI)
KEYS = ["r11", "r12", "r21", "r22", "r31", "r32", "r33", "r34", "r41", "r42", "r43", "r44", "r51"]


def uv_sample_fwd(layers, grid, out=None, clamp=CLAMP):
    ls = [l.detach().clamp(*clamp) for l in layers]
    ys = [F.grid_sample(l.unsqueeze(0), grid.unsqueeze(0), mode="bilinear", padding_mode="border",
                        align_corners=True)[0] for l in ls]
    return torch.stack(ys).sum(0)


def uv_scatter_bwd(grad_layers, grid, grad_out, hook0=None, hook1=None):
    g = grad_out.clone()
    if hook0 is not None:
        g = g * hook0.reshape(1, *g.shape[1:])
    if hook1 is not None:
        g = g * hook1.reshape(1, *g.shape[1:])
    for gl in grad_layers:
        with torch.enable_grad():            # the product calls us from inside an autograd.Function's backward
            z = torch.zeros_like(gl).requires_grad_(True)
            y = F.grid_sample(z.unsqueeze(0), grid.unsqueeze(0), mode="bilinear", padding_mode="border",
                              align_corners=True)[0]
            (dz,) = torch.autograd.grad(y, z, g)
        gl += dz


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, reg_coef=0.0, grad_scale=1.0,
              clamp=CLAMP):
    x = param.clamp(*clamp)
    g = grad * grad_scale + reg_coef * x
    exp_avg.lerp_(g, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = exp_avg_sq.sqrt() / (bc2 ** 0.5) + eps
    param.copy_(x - (lr / bc1) * exp_avg / denom)
    grad.zero_()


def texreg_value(param, coef, out_accum, clamp=CLAMP):
    out_accum += coef * (param.clamp(*clamp) ** 2).sum()


def _segment_ends(param, seg_begin):
    return list(seg_begin[1:]) + [param.numel()]


def adam_step_segments(param, grad, exp_avg, exp_avg_sq, seg_begin, seg_reg_coef, lr, beta1, beta2, eps, step,
                       grad_scale=1.0, clamp=CLAMP):
    for a, b, c in zip(seg_begin, _segment_ends(param, seg_begin), seg_reg_coef):
        adam_step(param[a:b], grad[a:b], exp_avg[a:b], exp_avg_sq[a:b], lr, beta1, beta2, eps, step, reg_coef=c,
                  grad_scale=grad_scale, clamp=clamp)


def texreg_value_segments(param, seg_begin, seg_coef, out_accum, clamp=CLAMP):
    for a, b, c in zip(seg_begin, _segment_ends(param, seg_begin), seg_coef):
        texreg_value(param[a:b], c, out_accum, clamp)


# ---- view preparation: the numpy oracle behind the engine's view_* signatures (CPU host-logic tests only) -------------
def view_uv_to_grid(uv_hw3, want_mask=False, depth_at_uv=None):
    from oracle import view_prep_oracle as vo
    uv = uv_hw3.numpy()
    m = None
    if want_mask:
        m = (uv[:, :, 0] != 0) | (uv[:, :, 1] != 0)
        if depth_at_uv is not None:
            m = m & (depth_at_uv.numpy() > 0)
        m = torch.from_numpy(m)
    return torch.from_numpy(vo.uv_to_grid(uv)), m


def view_gather2d(src, ytab, xtab):
    return src[ytab.long()][:, xtab.long()]


def view_resize_linear(src, out_hw, tables=None, divisor=1.0):
    from oracle import view_prep_oracle as vo
    a = src.numpy() if src.dtype != torch.uint16 else src.view(torch.int16).numpy().view("uint16")
    if a.dtype == "uint16":
        a = a / float(divisor)
    return torch.from_numpy(vo.resize_linear_cv2(a, (int(out_hw[1]), int(out_hw[0]))).astype("float64"))


def view_depth_levels(depth, levels, min_depth, depth_is_f32=False):
    from oracle import view_prep_oracle as vo
    d = depth.numpy().astype("float32") if depth_is_f32 else depth.numpy()
    c, r, o, w = vo.depth_levels(d, levels, min_depth)
    return (torch.from_numpy(c), depth.float(), torch.from_numpy(r), torch.from_numpy(o), torch.from_numpy(w))


def view_rgb_pre(rgb):
    from oracle import view_prep_oracle as vo
    return torch.from_numpy(vo.rgb_pre(rgb.numpy()))


def view_angle_degrees(c):
    from oracle import view_prep_oracle as vo
    return torch.from_numpy(vo.angle_degrees(c.numpy()))


def view_erode3x3(x):
    k = torch.ones(1, 1, 3, 3)
    m = torch.clamp(F.conv2d(x.reshape(1, 1, x.shape[-2], x.shape[-1]), k, padding=1) / 9, 0, 1).reshape(x.shape)
    return x * (m == 1)


def view_level_masks(mask, rounded, other, interp_w, num_levels):
    """model/model.py:210-239 with the reference's own torch ops (the emulation doubles as the K5 kernels' reference)."""
    H, W = mask.shape[-2], mask.shape[-1]
    mask_f = mask.reshape(1, 1, H, W).float()
    r, o, w = rounded.reshape(1, 1, H, W), other.reshape(1, 1, H, W), interp_w.reshape(1, 1, H, W).float()
    lm, lw = [], []
    for i in range(num_levels):
        lm.append(orc.erode((((r == i) + (o == i)).float()) * mask_f)[0, 0])
        m1 = orc.erode((r == i) * mask_f) * w
        m2 = orc.erode((o == i) * mask_f) * (1 - w)
        lw.append((m1 + m2)[0, 0])
    return torch.stack(lm), torch.stack(lw)


def view_level_plan(src_mask, src_weight, angle_guidance, angle_degrees, threshold, level_hw, layer_hw, counts):
    """model.py:199,219,238,253-254 + cs:161,172-185 with torch ops; same outputs as engine.view_level_plan."""
    Hr, Wr = src_mask.shape[-2], src_mask.shape[-1]
    H, W = level_hw
    up = lambda t, mode: F.interpolate(t.reshape(1, 1, Hr, Wr).float(), (H, W), mode=mode)
    M = (up(src_mask, "nearest") > 0).float()
    out = {"hook0": up(angle_guidance, "bilinear").reshape(-1).contiguous() if angle_guidance is not None else None,
           "hook1": up(src_weight, "nearest").reshape(-1).contiguous() if src_weight is not None else None, "layers": []}
    vals = [int(M.sum())]
    split = angle_degrees is not None
    if split:
        passed = up(angle_degrees, "bilinear") < threshold
    for (h, w) in layer_hw:
        rec = {"mask": F.interpolate(M, (h, w), mode="nearest").reshape(-1).contiguous()}
        n = [int(rec["mask"].sum()), 0, 0]
        if split:
            rec["mask_pass"] = F.interpolate(M * passed, (h, w), mode="nearest").reshape(-1).contiguous()
            rec["mask_fail"] = F.interpolate(M * (~passed), (h, w), mode="nearest").reshape(-1).contiguous()
            n[1], n[2] = int(rec["mask_pass"].sum()), int(rec["mask_fail"].sum())
        out["layers"].append(rec)
        vals += n
    counts[:len(vals)] = torch.tensor(vals, dtype=counts.dtype)
    return out


def unit_gram(impl, f, rowmask, inv_n):
    fl = f.reshape(f.shape[0], -1)
    if rowmask is not None:
        fl = fl * rowmask.reshape(1, -1)
    return fl @ fl.t() * inv_n


class VGGEngine:
    def __init__(self, state_dict, conv_impl=None, gram_impl=None):
        self.sd = {k: v.detach().cpu().float() for k, v in state_dict.items()}
        self.slots = []
        self.device = torch.device("cpu")

    def close(self):
        pass

    def begin(self, H, W):
        for i, s in enumerate(self.slots):
            if s["size"] == (H, W):
                return i
        self.slots.append({"size": (H, W)})
        return len(self.slots) - 1

    def release_slots(self):
        self.slots = []

    def forward(self, slot, image, last_conv, keep=None):
        s = self.slots[slot]
        with torch.enable_grad():            # the product may call us from inside an autograd.Function (no-grad mode)
            img = image.detach().clone().requires_grad_(True)
            feats = orc.vgg_forward(self.sd, img.unsqueeze(0), KEYS[:last_conv + 1], as_written=False)
        s.update(img=img, feats=feats, pend={}, last=last_conv)

    def feature_shape(self, slot, conv):
        f = self.slots[slot]["feats"][KEYS[conv]]
        return f.shape[1], f.shape[2], f.shape[3]

    def feature(self, slot, conv):
        return self.slots[slot]["feats"][KEYS[conv]][0].detach().clone()

    def feature_nhwc(self, slot, conv, out=None):
        f = self.feature(slot, conv)
        return f.permute(1, 2, 0).reshape(-1, f.shape[0]).contiguous()

    def features(self, image, keys):
        img = image[0] if image.dim() == 4 else image
        return {k: v.detach() for k, v in orc.vgg_forward(self.sd, img.unsqueeze(0), keys, as_written=False).items()}

    def _masked(self, slot, conv, rowmask):
        f = self.slots[slot]["feats"][KEYS[conv]][0].detach()
        fl = f.reshape(f.shape[0], -1)
        return fl * rowmask.reshape(1, -1) if rowmask is not None else fl

    def gram(self, slot, conv, rowmask, inv_n):
        fl = self._masked(slot, conv, rowmask)
        return fl @ fl.t() * inv_n

    def style_term(self, slot, conv, rowmask, inv_n, target0, coef0, target1, coef1, loss_accum, prev_sum=None,
                   avg_len=1.0, gram_out=None):
        fl = self._masked(slot, conv, rowmask)
        G = fl @ fl.t() * inv_n
        if gram_out is not None:
            gram_out.copy_(G)
        ghat = G if prev_sum is None else (G + prev_sum) / avg_len
        ln = avg_len if prev_sum is not None else 1.0
        C = G.shape[0]
        d = coef0 * (ghat - target0)
        loss_accum += coef0 * ((ghat - target0) ** 2).mean()
        if target1 is not None:
            d = d + coef1 * (ghat - target1)
            loss_accum += coef1 * ((ghat - target1) ** 2).mean()
        if inv_n == 0:
            return
        bmat = (2 * inv_n / ln) * (2 * d / (C * C))
        df = (bmat @ fl)                      # Bmat symmetric: dF[c][p] = sum_k B[c][k] Fm[k][p]
        s = self.slots[slot]
        s["pend"][conv] = s["pend"].get(conv, 0) + df.reshape(self.slots[slot]["feats"][KEYS[conv]][0].shape)

    def content_term(self, slot, conv, target_nhwc, rowmask, coef_loss, coef_grad, loss_accum):
        f = self.slots[slot]["feats"][KEYS[conv]][0].detach()
        t = target_nhwc.reshape(f.shape[1], f.shape[2], f.shape[0]).permute(2, 0, 1)
        m = rowmask.reshape(1, f.shape[1], f.shape[2])
        d = (f - t) * m
        loss_accum += coef_loss * (d ** 2).sum()
        s = self.slots[slot]
        s["pend"][conv] = s["pend"].get(conv, 0) + coef_grad * d

    def backward(self, slot, H, W, out=None):
        s = self.slots[slot]
        if not s["pend"]:
            return torch.zeros(3, H, W)
        with torch.enable_grad():
            total = sum((s["feats"][KEYS[c]][0] * g).sum() for c, g in s["pend"].items())
            (gi,) = torch.autograd.grad(total, s["img"], retain_graph=True)
        s["pend"] = {}
        return gi


def install(monkeypatch):
    """Route the product's engine calls to the emulation (tests only)."""
    from stylemesh_b200 import engine
    from stylemesh_b200.model import model as pm
    for name in ["uv_sample_fwd", "uv_scatter_bwd", "adam_step", "texreg_value", "adam_step_segments",
                 "texreg_value_segments", "unit_gram", "VGGEngine", "view_uv_to_grid", "view_gather2d",
                 "view_resize_linear", "view_depth_levels", "view_rgb_pre", "view_angle_degrees", "view_erode3x3",
                 "view_level_masks", "view_level_plan"]:
        monkeypatch.setattr(engine, name, globals()[name])
    monkeypatch.setattr(engine, "require_cuda_device", lambda dev: None)
