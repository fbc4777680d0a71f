"""Host-side logic of the product (loss plan, per-term coefficients, hooks, fused flat state, optimizer, the
no-op backward contract) checked on CPU against the reference's golden fixtures by swapping the C-ABI engine for
tests/fake_engine.py (an emulation of the ABI's semantics).  The kernels themselves are checked on the GPU."""
import os

import pytest
import torch

import fake_engine
from make_golden import build_inputs, golden_case_specs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = list(golden_case_specs().keys())


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


@pytest.mark.parametrize("case", CASES)
def test_pipeline_host_logic_against_fixture(case, monkeypatch, tmp_path):
    fake_engine.install(monkeypatch)
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    from oracle import stylemesh_oracle as orc
    gold = torch.load(os.path.join(GOLDEN_DIR, f"{case}.pt"), weights_only=False)
    spec = gold["spec"]
    preset, sd, layers, view, style, hierarchical = build_inputs(spec)
    vgg_path = os.path.join(tmp_path, "vgg.pth")
    torch.save(sd, vgg_path)
    W, H = spec["tex_size"]
    mdl = TextureOptimizationStyleTransferPipeline(
        W, H, hierarchical_texture=hierarchical, hierarchical_layers=len(layers), random_texture_init=True,
        style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
        use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
        style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
        angle_threshold=preset["angle_threshold"], learning_rate=spec["learning_rate"],
        loss_weights=dict(preset["loss_weights"]), save_texture=False)
    mods = list(mdl.texture.layers) if hierarchical else [mdl.texture]

    def reset():
        with torch.no_grad():
            for m, t in zip(mods, layers):
                m.data.copy_(t)

    reset()
    batch = view.as_batch()
    out = mdl.training_step(batch, 0)
    buf = mdl._loss_buf
    got = {"style": float(buf[0]), "content": float(buf[1]), "tex_reg": float(buf[2]), "total": float(buf[3])}
    for k, v in gold["loss0"].items():
        assert rel(got[k], v) < 1e-4 or abs(got[k] - v) < 1e-6, (k, got[k], v)
    assert out["loss"].requires_grad
    out["loss"].backward()                                   # must be a no-op, not an error
    lam = float(mdl.loss_weights.get("tex_reg", 0.0))
    for l, (g, gg) in enumerate(zip(mdl._grad_tensors(), gold["grad0"])):
        assert mods[l].data.grad.data_ptr() == g.data_ptr()  # .grad is a view of the flat buffer
        reg = 0.0
        if lam > 0 and hierarchical:
            x = mods[l].data.detach().clamp(orc.CLAMP_LO, orc.CLAMP_HI)
            reg = lam * mdl.tex_reg_weights[l] * 2.0 * x / x.numel()
        assert (g + reg - gg).norm() <= 1e-4 * gg.norm() + 1e-12, (case, l)

    # ---- teacher-forced trajectory: every step starts from the REFERENCE's parameters and Adam moments ----
    # (free-running comparison is chaotic: ReLU / max-pool mask flips amplify 1e-7 differences ~100x per step even
    #  between two fp32 CPU implementations, see DESIGN.md "parity definition")
    mdl._fused["grad"].zero_()
    if case == "dip":
        mdl.vgg_loss.gram_cache = {k: [] for k in mdl.vgg_loss.style_layers}
    (opt,), (sched,) = mdl.configure_optimizers()
    st = mdl._ensure_fused_state()
    for i in range(spec["steps"]):
        prev = gold["states"][i - 1] if i > 0 else None
        with torch.no_grad():
            for l, (m, (a, b)) in enumerate(zip(mods, st["spans"])):
                m.data.copy_((prev["params"][l] if prev else layers[l]))
                st["exp_avg"][a:b].copy_((prev["exp_avg"][l] if prev else torch.zeros_like(layers[l])).reshape(-1))
                st["exp_avg_sq"][a:b].copy_((prev["exp_avg_sq"][l] if prev else torch.zeros_like(layers[l])).reshape(-1))
        opt._steps = i
        opt.zero_grad()
        res = mdl.training_step(batch, i)
        res["loss"].backward()
        opt.step()
        buf = mdl._loss_buf
        got = {"style": float(buf[0]), "content": float(buf[1]), "tex_reg": float(buf[2]), "total": float(buf[3])}
        for k, v in gold["traj"][i].items():
            assert rel(got[k], v) < 1e-3 or abs(got[k] - v) < 1e-6, (i, k, got[k], v)
        for l, m in enumerate(mods):
            assert_final_texels(m.data.detach(), gold["states"][i]["params"][l], gold["states"][i]["exp_avg"][l])
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - spec["learning_rate"]) < 1e-12      # StepLR(step_size=30): unchanged


def test_module_autograd_surface_on_fake_engine(monkeypatch, tmp_path):
    """ContentAndStyleLoss.forward(...) returns autograd-connected losses whose backward reproduces the oracle."""
    fake_engine.install(monkeypatch)
    from stylemesh_b200.model.losses.content_and_style_losses import ContentAndStyleLoss
    from oracle import stylemesh_oracle as orc
    spec = golden_case_specs()["with_angle"]
    preset, sd, layers, view, style, _ = build_inputs(spec)
    vgg_path = os.path.join(tmp_path, "vgg.pth")
    torch.save(sd, vgg_path)
    mod = ContentAndStyleLoss(vgg_path, style_weights=list(preset["style_weights"]),
                              angle_threshold=preset["angle_threshold"], style_pyramid_mode="multi")
    mod.set_style_image(style.unsqueeze(0))
    pred = torch.rand(1, 3, 48, 64) * 100 - 50
    mask = torch.ones(1, 1, 48, 64); mask[..., 50:] = 0
    p1 = pred.clone().requires_grad_(True)
    s, c, _ = mod([p1], view.rgb, [mask], view.angle_degrees)
    (1e-4 * s + 70 * c).backward()
    loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                  angle_threshold=preset["angle_threshold"], style_pyramid_mode="multi",
                                  as_written=False)
    loss.set_style_image(style.unsqueeze(0))
    p2 = pred.clone().requires_grad_(True)
    os_, oc_ = loss.loss([p2], view.rgb, [mask], view.angle_degrees)
    (1e-4 * os_ + 70 * oc_).backward()
    assert rel(float(s), float(os_)) < 1e-4 and rel(float(c), float(oc_)) < 1e-4
    assert (p1.grad - p2.grad).norm() <= 1e-4 * p2.grad.norm()


def test_packed_batch_round_trip():
    """staging.PackedBatch: the collated byte buffer reproduces every tensor of the 13-tuple (dtype, shape, bytes),
    the dataset index stays a host tensor, offsets are 256-byte aligned."""
    from stylemesh_b200 import staging, synthetic as syn
    v = syn.make_view(1003, (32, 48), [(32, 48), (48, 64)])
    batch = v.as_batch()
    p = staging.PackedBatch(batch, pin=False)
    assert p.nbytes % 256 == 0 and p.payload_bytes == v.h2d_bytes() + 2 * 16 * 4      # + the two 4x4 camera matrices
    back = p.views_of(p.host.clone())
    assert isinstance(back, tuple) and len(back) == 13 and isinstance(back[9], list)
    assert back[8] is batch[8]
    flat_a, flat_b = [], []
    staging._flatten(batch, flat_a, [])
    staging._flatten(back, flat_b, [])
    assert len(flat_a) == len(flat_b) == 14
    for a, b in zip(flat_a, flat_b):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    for m in p.meta:
        assert m is None or m[0] % 256 == 0


def _loss_modules(tmp_path, preset_name, gram_mode, pyramid_mode):
    from stylemesh_b200.model.losses.content_and_style_losses import ContentAndStyleLoss
    from oracle import stylemesh_oracle as orc
    spec = golden_case_specs()[preset_name]
    preset, sd, layers, view, style, _ = build_inputs(spec)
    vgg_path = os.path.join(tmp_path, "vgg.pth")
    torch.save(sd, vgg_path)
    mod = ContentAndStyleLoss(vgg_path, style_weights=list(preset["style_weights"]),
                              angle_threshold=preset["angle_threshold"], style_pyramid_mode=pyramid_mode,
                              gram_mode=gram_mode)
    mod.set_style_image(style.unsqueeze(0))
    loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                  angle_threshold=preset["angle_threshold"], style_pyramid_mode=pyramid_mode,
                                  gram_mode=gram_mode, as_written=False)
    loss.set_style_image(style.unsqueeze(0))
    return mod, loss, view


def test_autograd_average_gram_mode_two_levels_uses_the_forward_histories(monkeypatch, tmp_path):
    """gram_mode='average' with a 2-level pyramid (cs:319-323: the cache is shared across levels, so level 1 averages
    with level 0's Gram of the SAME call): losses and gradients of three consecutive calls equal the oracle's, also
    when another forward (a validation pass) runs between a forward and its backward."""
    fake_engine.install(monkeypatch)
    mod, loss, view = _loss_modules(tmp_path, "only2D", "average", "single")
    g = torch.Generator().manual_seed(5)
    masks = [torch.ones(1, 1, 48, 64), torch.ones(1, 1, 72, 96)]
    masks[0][..., 50:] = 0
    masks[1][..., :20] = 0
    for call in range(3):
        preds = [torch.rand(1, 3, 48, 64, generator=g) * 100 - 50, torch.rand(1, 3, 72, 96, generator=g) * 100 - 50]
        p1 = [p.clone().requires_grad_(True) for p in preds]
        s, c, _ = mod(p1, view.rgb, masks, view.angle_degrees)
        if call == 1:                                    # an unrelated forward before the backward (validation)
            with torch.no_grad():
                saved = {k: list(v) for k, v in mod.gram_cache.items()}
                mod([p.detach() for p in preds], view.rgb, masks, view.angle_degrees)
                mod.gram_cache = saved                   # (the oracle does not see this extra call)
        (1e-4 * s + 70 * c).backward()
        p2 = [p.clone().requires_grad_(True) for p in preds]
        os_, oc_ = loss.loss(p2, view.rgb, masks, view.angle_degrees)
        (1e-4 * os_ + 70 * oc_).backward()
        assert rel(float(s), float(os_)) < 1e-4 and rel(float(c), float(oc_)) < 1e-4, call
        for a, b in zip(p1, p2):
            assert (a.grad - b.grad).norm() <= 1e-4 * b.grad.norm(), call


def test_soft_masks_weight_the_levels_like_the_reference(monkeypatch, tmp_path):
    """cs:181: the level factor is mean(nearest(mask)) of the RAW mask values; `> 0` only selects the pixels (cs:137)."""
    fake_engine.install(monkeypatch)
    mod, loss, view = _loss_modules(tmp_path, "only2D", "current", "single")
    g = torch.Generator().manual_seed(9)
    masks = [torch.rand(1, 1, 48, 64, generator=g) * 0.5, torch.rand(1, 1, 72, 96, generator=g) * 2.0]
    masks[0][..., 40:] = 0
    preds = [torch.rand(1, 3, 48, 64, generator=g) * 100 - 50, torch.rand(1, 3, 72, 96, generator=g) * 100 - 50]
    s, c, info = mod(preds, view.rgb, masks, view.angle_degrees)
    os_, oc_ = loss.loss(preds, view.rgb, masks, view.angle_degrees)
    assert rel(float(s), float(os_)) < 1e-4 and rel(float(c), float(oc_)) < 1e-4


def test_checkpoint_round_trip_restores_texels_moments_step_and_lr(monkeypatch, tmp_path):
    """Trainer.save_checkpoint / load_checkpoint (Lightning's implicit ModelCheckpoint + --resume_from_checkpoint):
    texture layers, Adam moments + step in torch.optim.Adam's layout, StepLR state, epoch, global_step."""
    fake_engine.install(monkeypatch)
    from stylemesh_b200.lightning_shim import Trainer
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    spec = golden_case_specs()["only2D"]
    preset, sd, layers, view, style, hierarchical = build_inputs(spec)
    vgg_path = os.path.join(tmp_path, "vgg.pth")
    torch.save(sd, vgg_path)

    def make():
        return TextureOptimizationStyleTransferPipeline(
            64, 64, hierarchical_texture=True, hierarchical_layers=len(layers), random_texture_init=True,
            style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
            use_angle_weight=False, use_depth_scaling=False, learning_rate=1.0, decay_gamma=0.5, decay_step_size=1,
            loss_weights=dict(preset["loss_weights"]), save_texture=False)

    a = make()
    (opt,), (sched,) = a.configure_optimizers()
    for i in range(3):
        a.training_step(view.as_batch(), i)["loss"].backward()
        opt.step()
    sched.step()
    tr = Trainer(default_root_dir=str(tmp_path))
    tr.global_step = 3
    path = tr.save_checkpoint(a, opt, [sched], epoch=0)
    assert os.path.basename(path) == "epoch=0-step=3.ckpt"
    ck = torch.load(path, weights_only=False)
    ref_adam = torch.optim.Adam([torch.nn.Parameter(t.clone()) for t in layers], lr=1.0)
    ref_adam.load_state_dict(ck["optimizer_states"][0])               # the reference's optimizer accepts our state
    assert int(ref_adam.state_dict()["state"][0]["step"]) == 3

    b = make()
    (opt_b,), (sched_b,) = b.configure_optimizers()
    tr2 = Trainer(default_root_dir=str(tmp_path), resume_from_checkpoint=path)
    tr2.load_checkpoint(path, b, opt_b, [sched_b])
    assert tr2.start_epoch == 1 and tr2.global_step == 3 and opt_b._steps == 3
    assert abs(opt_b.param_groups[0]["lr"] - 0.5) < 1e-12
    for ma, mb in zip(a.texture.layers, b.texture.layers):
        assert torch.equal(ma.data, mb.data)
    sa, sb = a._ensure_fused_state(), b._ensure_fused_state()
    assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"])
    # the next step of both is identical
    for m, o in ((a, opt), (b, opt_b)):
        m.training_step(view.as_batch(), 3)["loss"].backward()
        o.step()
    for ma, mb in zip(a.texture.layers, b.texture.layers):
        assert torch.equal(ma.data, mb.data)


def test_fused_pipeline_average_gram_mode_with_a_multi_level_pyramid(monkeypatch, tmp_path):
    """gram_mode='average' (cs:319-323) on a 3-level pyramid with the angle split ('multi'): the Gram cache is shared by
    the levels of a step and carried across steps; the fused training path must follow the oracle over several steps
    (losses and dense gradients, teacher-forced on identical texels)."""
    fake_engine.install(monkeypatch)
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    from oracle import stylemesh_oracle as orc
    spec = golden_case_specs()["with_angle_and_depth"]
    preset, sd, layers, view, style, hierarchical = build_inputs(spec)
    vgg_path = os.path.join(tmp_path, "vgg.pth")
    torch.save(sd, vgg_path)
    W, H = spec["tex_size"]
    mdl = TextureOptimizationStyleTransferPipeline(
        W, H, hierarchical_texture=True, hierarchical_layers=len(layers), random_texture_init=True,
        style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
        use_angle_weight=True, use_depth_scaling=True, style_pyramid_mode="multi", gram_mode="average",
        angle_threshold=preset["angle_threshold"], learning_rate=1.0, loss_weights=dict(preset["loss_weights"]),
        save_texture=False)
    with torch.no_grad():
        for m, t in zip(mdl.texture.layers, layers):
            m.data.copy_(t)
    loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                  angle_threshold=preset["angle_threshold"], style_pyramid_mode="multi",
                                  gram_mode="average", as_written=False)
    loss.set_style_image(style.unsqueeze(0))
    cfg = orc.OracleConfig(use_angle_weight=True, use_depth_scaling=True, loss_weights=dict(preset["loss_weights"]),
                           hierarchical=True, learning_rate=1.0)
    pipe = orc.OraclePipeline(layers, loss, cfg)
    (opt,), _ = mdl.configure_optimizers()
    batch = view.as_batch()
    lam = float(mdl.loss_weights["tex_reg"])
    for step in range(3):
        with torch.no_grad():
            for t, m in zip(pipe.layers, mdl.texture.layers):
                t.copy_(m.data)
        want, want_grads = pipe.grads(batch)
        out = mdl.training_step(batch, step)
        buf = mdl._loss_buf
        for k, i in (("style", 0), ("content", 1), ("tex_reg", 2), ("total", 3)):
            assert rel(float(buf[i]), want[k]) < 1e-4, (step, k, float(buf[i]), want[k])
        for l, (g, gg) in enumerate(zip(mdl._grad_tensors(), want_grads)):
            x = pipe.layers[l].detach().clamp(orc.CLAMP_LO, orc.CLAMP_HI)
            data_want = gg - lam * mdl.tex_reg_weights[l] * 2.0 * x / x.numel()
            assert (g - data_want).norm() <= 1e-4 * data_want.norm() + 1e-12, (step, l)
        out["loss"].backward()
        opt.step()
    assert all(len(v) == 9 for v in mdl.vgg_loss.gram_cache.values())        # 3 levels x 3 steps, capped at 10
