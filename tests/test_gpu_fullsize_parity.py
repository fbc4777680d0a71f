"""Teacher-forced parity AT THE SIZES BASELINE.json NAMES, CUDA path vs the CPU oracle on the same seeded inputs:

  C2  only2D               640x480 view, 2048^2 x 4-layer texture, style 768x970
  C3  with_angle_and_depth rgb 256x341, ScanNet UV pyramid {256x341, 432x576, 608x811, 784x1045}, 2048^2 x 4, 'multi' mode
      (/root/reference/scripts/train/optimize_texture_scannet_with_angle_and_depth.sh:3-28, model/model.py:210-251)
  C4  with_angle_and_depth rgb 256x320, Matterport pyramid {256x320 ... 784x980}, 4096^2 x 4-layer texture

The oracle needs 0.6 s (C2) to a few seconds (C3/C4) per step on the box's host cores.  At these sizes the kernels run
the code paths the 48x64 fixtures never reach: stream-K splits, resident B tiles, 4-buffer TMEM at BN=64, the 146-way
Gram split, ragged TMA clipping on 341/811/1045-pixel rows.

Bars: loss terms <= 1e-3 relative (measured <= 6e-5); texels after one teacher-forced Adam step by distribution (DESIGN
§5).  The dense texture gradient is judged against the EXACT gradient: the same oracle run in float64.  The gradient
of a ReLU / max-pool network is discontinuous in the activations, so the reference's own fp32 arithmetic is 0.5 % (C2)
to 1.4 % (C4, 4096^2 texture: fewer pixels averaged per texel) away from the exact gradient in relative L2
(profiles/r02_gradient_noise_floor.md).  The CUDA path must be as close to the exact gradient as the fp32 reference
arithmetic is, within a factor of two: err(ours, f64) <= max(1e-2, 2 x err(fp32 oracle, f64)); measured 1.3 - 1.8 x.

Unit shapes: the convs / Grams of the benchmark view at the layers where igemm_ph<64> and the r11/r21 Gram kernels run
(64x480x640, 128x240x320) against torch fp32/fp64 CPU.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import stylemesh_oracle as orc

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-3
GRAD_TOL = 1e-2

SCANNET_PYRAMID = [(256, 341), (432, 576), (608, 811), (784, 1045)]       # scripts/scannet/render_uvs.py:79-80,126-130
MATTERPORT_PYRAMID = [(256, 320), (432, 540), (608, 760), (784, 980)]

FULL_CASES = {
    "C2_only2D_640x480_tex2048": dict(preset="only2D", rgb=(480, 640), levels=[(480, 640)], tex=2048),
    "C3_angle_depth_scannet_tex2048": dict(preset="with_angle_and_depth", rgb=(256, 341), levels=SCANNET_PYRAMID,
                                           tex=2048),
    "C4_angle_depth_matterport_tex4096": dict(preset="with_angle_and_depth", rgb=(256, 320), levels=MATTERPORT_PYRAMID,
                                              tex=4096),
}


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def _log(record):
    path = os.environ.get("SMB_PARITY_LOG")
    if path:
        with open(path, "a") as fh:
            fh.write(json.dumps(record) + "\n")


def _oracle_f64_grads(case, view_seed=1000):
    """the same oracle, same seeded inputs (generated in fp32, then widened), evaluated in float64: the exact gradient"""
    from stylemesh_b200 import synthetic as syn
    preset = syn.PRESETS[case["preset"]]
    f64 = torch.float64
    sd = {k: v.to(f64) for k, v in syn.make_vgg_state_dict(0, bias_scale=0.05).items()}
    layers = [t.to(f64) for t in syn.make_texture_layers(11, case["tex"], case["tex"], preset["hierarchical_layers"])]
    view = syn.make_view(view_seed, case["rgb"], case["levels"])
    style = syn.make_style_image(7, 768, 970).to(f64)
    widen = lambda x: x.to(f64) if isinstance(x, torch.Tensor) and x.is_floating_point() else x
    batch = tuple([widen(u) for u in x] if isinstance(x, list) else widen(x) for x in view.as_batch())
    old = torch.get_default_dtype()
    torch.set_default_dtype(f64)
    try:
        loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                      angle_threshold=preset["angle_threshold"],
                                      style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
                                      as_written=False)
        loss.set_style_image(style.unsqueeze(0))
        cfg = orc.OracleConfig(use_angle_weight=preset["use_angle_weight"],
                               use_depth_scaling=preset["use_depth_scaling"],
                               loss_weights=dict(preset["loss_weights"]), hierarchical=True, learning_rate=1.0)
        return orc.OraclePipeline(layers, loss, cfg).grads(batch)
    finally:
        torch.set_default_dtype(old)


def _build(case, tmp_path, view_seed=1000):
    from stylemesh_b200 import synthetic as syn
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    os.environ.pop("SMB_CONV_IMPL", None)
    os.environ.pop("SMB_GRAM_IMPL", None)
    preset = syn.PRESETS[case["preset"]]
    sd = syn.make_vgg_state_dict(0, bias_scale=0.05)
    nl = preset["hierarchical_layers"]
    layers = syn.make_texture_layers(11, case["tex"], case["tex"], nl)
    view = syn.make_view(view_seed, case["rgb"], case["levels"])
    style = syn.make_style_image(7, 768, 970)
    vgg_path = os.path.join(str(tmp_path), "vgg.pth")
    torch.save(sd, vgg_path)
    mdl = TextureOptimizationStyleTransferPipeline(
        case["tex"], case["tex"], hierarchical_texture=True, hierarchical_layers=nl, random_texture_init=True,
        style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
        use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
        style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
        angle_threshold=preset["angle_threshold"], learning_rate=1.0, loss_weights=dict(preset["loss_weights"]),
        save_texture=False)
    mdl.cuda()
    with torch.no_grad():
        for m, t in zip(mdl.texture.layers, layers):
            m.data.copy_(t.cuda())
    loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                  angle_threshold=preset["angle_threshold"],
                                  style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
                                  as_written=False)
    loss.set_style_image(style.unsqueeze(0))
    cfg = orc.OracleConfig(use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
                           loss_weights=dict(preset["loss_weights"]), hierarchical=True, learning_rate=1.0)
    pipe = orc.OraclePipeline(layers, loss, cfg)
    return mdl, pipe, view, preset


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_full_size_step_matches_oracle(name, tmp_path):
    case = FULL_CASES[name]
    mdl, pipe, view, preset = _build(case, tmp_path)
    batch = view.to("cuda").as_batch()
    want_loss, want_grads = pipe.grads(view.as_batch())

    (opt,), _ = mdl.configure_optimizers()
    out = mdl.training_step(batch, 0)
    buf = mdl._loss_buf.cpu()
    got = {"style": float(buf[0]), "content": float(buf[1]), "tex_reg": float(buf[2]), "total": float(buf[3])}
    _log({"kind": "fullsize_loss", "case": name, "rel": {k: rel(got[k], want_loss[k]) for k in got}, "got": got})
    for k in got:
        assert rel(got[k], want_loss[k]) < LOSS_TOL or abs(got[k] - want_loss[k]) < 1e-6, (k, got[k], want_loss[k])

    exact_loss, exact_grads = _oracle_f64_grads(case)
    for k in got:
        assert rel(got[k], exact_loss[k]) < LOSS_TOL or abs(got[k] - exact_loss[k]) < 1e-6, (k, got[k], exact_loss[k])
    lam = float(mdl.loss_weights.get("tex_reg", 0.0))
    for l, (g, g32, g64) in enumerate(zip([g.cpu() for g in mdl._grad_tensors()], want_grads, exact_grads)):
        x = pipe.layers[l].detach().clamp(orc.CLAMP_LO, orc.CLAMP_HI).double()
        reg = lam * mdl.tex_reg_weights[l] * 2.0 * x / x.numel()      # added by the fused Adam kernel on our side
        exact = g64 - reg                                              # the data term of the exact gradient
        n = exact.norm().item()
        e_ours = (g.double() - exact).norm().item() / n
        e_ref32 = (g32.double() - reg - exact).norm().item() / n
        e_ours_vs32 = (g.double() - (g32.double() - reg)).norm().item() / (g32.double() - reg).norm().item()
        _log({"kind": "fullsize_grad", "case": name, "layer": l, "rel_l2_ours_vs_f64": e_ours,
              "rel_l2_fp32oracle_vs_f64": e_ref32, "rel_l2_ours_vs_fp32oracle": e_ours_vs32})
        assert e_ours <= max(GRAD_TOL, 2.0 * e_ref32), (name, l, e_ours, e_ref32, e_ours_vs32)

    # ---- one teacher-forced Adam step from identical parameters and zero moments ----
    out["loss"].backward()
    opt.step()
    pipe.opt.step()                        # pipe.grads() left .grad populated with the oracle's gradient
    for l, (m, t) in enumerate(zip(mdl.texture.layers, pipe.layers)):
        a, b = m.data.detach().cpu(), t.detach()
        d = (a - b).abs()
        off = (d > 1e-3 * b.abs().clamp_min(1.0)).float().mean().item()
        flipped = (d > 0.1).float().mean().item()
        _log({"kind": "fullsize_texels", "case": name, "layer": l, "frac_off_gt_1e-3": off, "frac_flipped": flipped,
              "median_abs": d.median().item()})
        assert off <= 5e-2 and flipped <= 5e-3 and d.median().item() <= 1e-5, (name, l, off, flipped, d.median().item())


def test_second_visit_uses_the_caches_and_still_matches_oracle(tmp_path):
    """cache_view_plans + cache_content_targets (replaces re-running cs:294 and model.py:188-257 for a
    repeated view, abstract_dataset.py:498-512): visit two views twice each, free-running; every visit - the cache
    fills on the first, the hits on the second - is compared with the oracle teacher-forced on our parameters."""
    case = dict(preset="with_angle_and_depth", rgb=(128, 171), levels=[(128, 171), (216, 288), (304, 405)], tex=512)
    mdl, pipe, view_a, preset = _build(case, tmp_path, view_seed=1000)
    from stylemesh_b200 import synthetic as syn
    view_b = syn.make_view(1001, case["rgb"], case["levels"])
    mdl.cache_view_plans = True
    mdl.vgg_loss.cache_content_targets = True
    (opt,), _ = mdl.configure_optimizers()
    views = [view_a, view_b]
    dev = [v.to("cuda").as_batch() for v in views]
    hits = []
    for visit, vi in enumerate([0, 1, 0, 1, 0]):
        with torch.no_grad():                       # teacher-force the oracle on OUR current texels
            for t, m in zip(pipe.layers, mdl.texture.layers):
                t.copy_(m.data.detach().cpu())
        want_loss, want_grads = pipe.grads(views[vi].as_batch())
        n_plans, n_tgts = len(mdl._plan_cache), len(mdl.vgg_loss._content_cache)
        out = mdl.training_step(dev[vi], visit)
        hits.append((len(mdl._plan_cache) == n_plans, len(mdl.vgg_loss._content_cache) == n_tgts))
        buf = mdl._loss_buf.cpu().clone()
        for k, i in (("style", 0), ("content", 1), ("tex_reg", 2), ("total", 3)):
            assert rel(float(buf[i]), want_loss[k]) < LOSS_TOL, (visit, k, float(buf[i]), want_loss[k])
        lam = float(mdl.loss_weights["tex_reg"])
        grads_cached = [g.clone() for g in mdl._grad_tensors()]
        for l, (g, gg) in enumerate(zip([g.cpu() for g in grads_cached], want_grads)):
            x = pipe.layers[l].detach().clamp(orc.CLAMP_LO, orc.CLAMP_HI)
            data_want = gg - lam * mdl.tex_reg_weights[l] * 2.0 * x / x.numel()
            # vs the fp32 oracle: bounded by the oracle's own gradient noise floor (profiles/r02_gradient_noise_floor.md)
            assert (g - data_want).norm() <= 3e-2 * data_want.norm() + 1e-12, (visit, l)
        # the same step recomputed with both caches OFF: everything up to the scatter is deterministic, the scatter
        # adds with float atomics -> equal to rounding
        mdl._fused["grad"].zero_()
        mdl.cache_view_plans, mdl.vgg_loss.cache_content_targets = False, False
        buf2 = mdl.fused_view_step(dev[vi]).cpu().clone()
        mdl.cache_view_plans, mdl.vgg_loss.cache_content_targets = True, True
        assert torch.allclose(buf, buf2, rtol=1e-5, atol=0), (visit, buf, buf2)
        for l, (a, b) in enumerate(zip(grads_cached, mdl._grad_tensors())):
            assert (a - b).norm() <= 1e-5 * b.norm() + 1e-20, (visit, l, float((a - b).norm() / b.norm()))
        out["loss"].backward()
        opt.step()
    assert hits == [(False, False), (False, False), (True, True), (True, True), (True, True)], hits


# ---------------------------------------------------------------------------------------------------------------
# unit shapes of the benchmark view
# ---------------------------------------------------------------------------------------------------------------
BENCH_CONV_SHAPES = [(64, 64, 480, 640),      # conv1_2: igemm_ph<64>, resident B tiles, 4 TMEM buffers
                     (64, 128, 240, 320),     # conv2_1 (Cin = Cout/2)
                     (128, 128, 240, 320),    # conv2_2
                     (128, 64, 240, 320)]     # conv2_1 data gradient shape (N = 64)


@pytest.mark.parametrize("cin,cout,h,w", BENCH_CONV_SHAPES)
def test_conv_forward_and_data_gradient_at_benchmark_shapes(cin, cout, h, w):
    from stylemesh_b200 import engine as eng
    g = torch.Generator().manual_seed(cin * 3 + cout + h)
    x = torch.randn(1, cin, h, w, generator=g) * 50
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, padding=1))[0]
    out = eng.unit_conv3x3(5, x[0].cuda(), wt, b, relu=True).cpu()
    r = float((out - ref).norm() / ref.norm())
    _log({"kind": "unit_conv_fwd", "shape": [cin, cout, h, w], "rel_l2": r})
    assert r < 5e-5, r
    dy = torch.randn(1, cout, h, w, generator=g)
    xg = x.clone().requires_grad_(True)
    F.conv2d(xg, wt, None, padding=1).backward(dy)
    dg = eng.unit_conv3x3(5, dy[0].cuda(), wt, None, relu=False, transpose_flip=True).cpu()
    r = float((dg - xg.grad[0]).norm() / xg.grad[0].norm())
    _log({"kind": "unit_conv_dgrad", "shape": [cin, cout, h, w], "rel_l2": r})
    assert r < 5e-5, r


@pytest.mark.parametrize("c,h,w", [(64, 480, 640), (128, 240, 320), (256, 120, 160)])
def test_masked_gram_at_benchmark_shapes(c, h, w):
    from stylemesh_b200 import engine as eng
    g = torch.Generator().manual_seed(c + h)
    f = F.relu(torch.randn(c, h, w, generator=g)) * 30
    mask = (torch.rand(h * w, generator=g) > 0.1).float()
    n = float(mask.sum())
    fm = (f.reshape(c, -1) * mask).double()
    ref = fm @ fm.t() / n
    out = eng.unit_gram(1, f.cuda(), mask.cuda(), 1.0 / n).cpu()
    r = float((out.double() - ref).norm() / ref.norm())
    _log({"kind": "unit_gram", "shape": [c, h, w], "rel_l2": r})
    assert r < 2e-5, r
