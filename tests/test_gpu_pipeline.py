"""End-to-end GPU parity of the hot path (through the reference-facing modules and the C-ABI) against
 (a) the golden fixtures produced by the real reference and (b) the CPU oracle on the same seeded inputs.
Tolerances follow BASELINE.json north_star: <= 1e-3 relative on loss values and on texture texels (norm-wise,
SURVEY §7b), gradients <= 1e-3 relative L2."""
import os

import pytest
import torch

from make_golden import build_inputs, golden_case_specs
from oracle import stylemesh_oracle as orc

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = list(golden_case_specs().keys())
IMPLS = [pytest.param("tc", id="tc"), pytest.param("simt", id="simt")]


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def assert_final_texels(got, want, grad0):
    """Adam with lr=1 (all reference scripts) turns a texel gradient g into an update g/(|g|+1e-8): texels whose
    gradient is ~eps amplify 1e-9 absolute noise into O(0.1) texel differences.  The 1e-3 bar is
    therefore asserted on texels with a non-negligible gradient; all texels must still agree to 5e-3."""
    well = grad0.abs() > 1e-5 * grad0.abs().max()
    d = got - want
    assert d[well].norm() <= 1e-3 * want[well].norm(), (float(d[well].norm()), float(want[well].norm()))
    assert d.norm() <= 5e-3 * want.norm(), (float(d.norm()), float(want.norm()))


def make_pipeline(spec, tmp_path, impl):
    os.environ["SMB_CONV_IMPL"] = impl
    os.environ["SMB_GRAM_IMPL"] = impl
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    preset, sd, layers, view, style, hierarchical = build_inputs(spec)
    vgg_path = os.path.join(tmp_path, "vgg_synth.pth")
    torch.save(sd, vgg_path)
    W, H = spec["tex_size"]
    mdl = TextureOptimizationStyleTransferPipeline(
        W, H, hierarchical_texture=hierarchical, hierarchical_layers=len(layers), random_texture_init=True,
        style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
        use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
        style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
        angle_threshold=preset["angle_threshold"], learning_rate=spec["learning_rate"], decay_gamma=0.1,
        decay_step_size=30, loss_weights=dict(preset["loss_weights"]), tex_reg_weights=None, save_texture=False)
    mdl.cuda()

    def reset_texture():
        mods = list(mdl.texture.layers) if hierarchical else [mdl.texture]
        with torch.no_grad():
            for m, t in zip(mods, layers):
                m.data.copy_(t.cuda())
        return mods

    mods = reset_texture()
    batch = view.to("cuda").as_batch()
    return mdl, mods, batch, reset_texture, (preset, sd, layers, view, style, hierarchical)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("case", CASES)
def test_step_matches_reference_fixture(case, impl, tmp_path):
    gold = torch.load(os.path.join(GOLDEN_DIR, f"{case}.pt"), weights_only=False)
    spec = gold["spec"]
    mdl, mods, batch, reset_texture, (_, _, layers, *_unused) = make_pipeline(spec, str(tmp_path), impl)

    # ---- style targets (Gram of the style pyramid) ----
    mdl._ensure_style_targets(batch[0])
    t = mdl.vgg_loss.style_targets
    assert (t[0][0].cpu() - gold["style_target_r11_l0"]).norm() <= 1e-4 * gold["style_target_r11_l0"].norm()
    for i, row in enumerate(gold["style_target_sums"]):
        for l, s in enumerate(row):
            assert rel(float(t[i][l].sum()), s) < 1e-4, (i, l)

    # ---- teacher-forced: loss terms and dense texture gradient ----
    out = mdl.training_step(batch, 0)
    buf = mdl._loss_buf.cpu()
    got = {"style": float(buf[0]), "content": float(buf[1]), "tex_reg": float(buf[2]), "total": float(buf[3])}
    for k, v in gold["loss0"].items():
        assert rel(got[k], v) < 1e-3 or abs(got[k] - v) < 1e-6, (k, got[k], v)
    assert rel(float(out["loss"]), gold["loss0"]["total"]) < 1e-3
    grads = [g.cpu() for g in mdl._grad_tensors()]
    lam = float(mdl.loss_weights.get("tex_reg", 0.0))
    for l, (g, gg) in enumerate(zip(grads, gold["grad0"])):
        # the fixture's gradient includes the regulariser term, which the fused Adam kernel adds itself
        reg = 0.0
        if lam > 0 and mdl.hierarchical_texture:
            x = mods[l].data.detach().cpu().clamp(orc.CLAMP_LO, orc.CLAMP_HI)
            reg = lam * mdl.tex_reg_weights[l] * 2.0 * x / x.numel()
        err = (g + reg - gg).norm().item()
        assert err <= 1e-3 * gg.norm().item() + 1e-12, (case, l, err, gg.norm().item())

    # ---- free-running trajectory: `steps` fused Adam steps from the same start ----
    # ---- teacher-forced trajectory: every step starts from the REFERENCE's parameters and Adam moments ----
    # (free-running comparison is chaotic: ReLU / max-pool mask flips amplify 1e-7 differences ~100x per step even
    #  between two fp32 CPU implementations, see DESIGN.md "parity definition")
    mdl._fused["grad"].zero_()
    if case == "dip":
        mdl.vgg_loss.gram_cache = {k: [] for k in mdl.vgg_loss.style_layers}
    (opt,), _ = mdl.configure_optimizers()
    st = mdl._ensure_fused_state()
    for i in range(spec["steps"]):
        prev = gold["states"][i - 1] if i > 0 else None
        with torch.no_grad():
            for l, (m, (a, b)) in enumerate(zip(mods, st["spans"])):
                m.data.copy_((prev["params"][l] if prev else layers[l]).cuda())
                st["exp_avg"][a:b].copy_((prev["exp_avg"][l] if prev else torch.zeros_like(layers[l])).reshape(-1).cuda())
                st["exp_avg_sq"][a:b].copy_((prev["exp_avg_sq"][l] if prev else torch.zeros_like(layers[l])).reshape(-1).cuda())
        opt._steps = i
        opt.zero_grad()
        res = mdl.training_step(batch, i)
        res["loss"].backward()
        opt.step()
        buf = mdl._loss_buf.cpu()
        got = {"style": float(buf[0]), "content": float(buf[1]), "tex_reg": float(buf[2]), "total": float(buf[3])}
        for k, v in gold["traj"][i].items():
            assert rel(got[k], v) < 1e-3 or abs(got[k] - v) < 1e-6, (i, k, got[k], v)
        for l, m in enumerate(mods):
            assert_final_texels(m.data.detach().cpu(), gold["states"][i]["params"][l], gold["states"][i]["exp_avg"][l])


def test_vgg_features_match_oracle(tmp_path):
    spec = golden_case_specs()["only2D"]
    mdl, _, batch, _, (preset, sd, *_rest) = make_pipeline(spec, str(tmp_path), "tc")
    keys = ["r11", "r21", "r31", "r41", "r42", "r51"]
    x = batch[0]
    got = mdl.vgg_loss.vgg(x, keys)
    want = orc.vgg_forward(sd, x.cpu(), keys, as_written=False)
    for k in keys:
        r = float((got[k].cpu() - want[k]).norm() / want[k].norm())
        assert r < 1e-4, (k, r)


def test_module_level_autograd_api_matches_oracle(tmp_path):
    """ContentAndStyleLoss.forward + texture.forward through torch autograd (the drop-in module surface)."""
    spec = golden_case_specs()["with_angle"]
    mdl, mods, batch, _, (preset, sd, layers, view, style, hierarchical) = make_pipeline(spec, str(tmp_path), "tc")
    mdl._ensure_style_targets(batch[0])
    pred = [mdl.texture(v) for v in batch[9]]
    mask = (torch.nn.functional.interpolate(batch[10].unsqueeze(1).float(), pred[-1].shape[2:], mode="nearest") > 0).float()
    style_l, content_l, info = mdl.vgg_loss(pred, batch[0], [mask], batch[12])
    total = 1e-4 * style_l + 70.0 * content_l
    for m in mods:
        m.data.grad = None
    total.backward()

    loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                  angle_threshold=preset["angle_threshold"],
                                  style_pyramid_mode=preset["style_pyramid_mode"], as_written=False)
    loss.set_style_image(style.unsqueeze(0))
    ol = [t.clone().requires_grad_(True) for t in layers]
    opred = [orc.texture_sample(ol, v) for v in view.uvs]
    os_, oc_ = loss.loss(opred, view.rgb, [mask.cpu()], view.angle_degrees)
    (1e-4 * os_ + 70.0 * oc_).backward()
    assert rel(float(style_l), float(os_)) < 1e-3 and rel(float(content_l), float(oc_)) < 1e-3
    for m, o in zip(mods, ol):
        assert (m.data.grad.cpu() - o.grad).norm() <= 1e-3 * o.grad.norm()
