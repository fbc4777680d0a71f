"""The strongest check of the drop-in boundary: the reference's OWN, UNMODIFIED LightningModule
(/root/reference/model/model.py, loaded from where it lies, never copied) runs on top of THIS repo's `model.texture.*`
and `model.losses.*` modules - its `from model.texture.texture import ...` / `from model.losses... import ...` lines
resolve to the repo's packages - and must reproduce the golden fixtures that the all-reference stack produced
(tests/golden/make_golden.py): loss terms, dense texture gradients through torch autograd, and the teacher-forced Adam
trajectory with the reference's torch.optim.Adam.

Runs only in the build container (needs /root/reference; the GPU box does not have it) with the engine emulated on
CPU (tests/fake_engine.py): what is under test is the module API - names, ctor kwargs, argument meaning, return types,
autograd connectivity - not the kernels (tests/test_gpu_pipeline.py covers those)."""
import importlib.util
import os
import sys

import pytest
import torch

import fake_engine
from make_golden import REFERENCE, _install_lightning_stub, build_inputs, golden_case_specs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_MODEL_PY = os.path.join(REFERENCE, "model", "model.py")

pytestmark = pytest.mark.skipif(not os.path.isfile(REF_MODEL_PY), reason="needs the reference checkout")


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def _load_reference_lightning_module():
    """exec the reference's model/model.py under a private module name; every `model.*` import inside it binds to the
    repo's shim packages (repo root precedes /root/reference on sys.path, and /root/reference is not on it at all)"""
    _install_lightning_stub()
    assert REFERENCE not in sys.path
    import model.texture.texture as tex_mod
    import model.losses.content_and_style_losses as loss_mod
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert os.path.abspath(tex_mod.__file__).startswith(repo) and os.path.abspath(loss_mod.__file__).startswith(repo)
    spec = importlib.util.spec_from_file_location("reference_model_py_under_test", REF_MODEL_PY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.ContentAndStyleLoss is loss_mod.ContentAndStyleLoss          # ours, not the reference's
    assert mod.HierarchicalNeuralTexture is tex_mod.HierarchicalNeuralTexture
    return mod


@pytest.mark.parametrize("case", ["only2D", "with_angle", "with_angle_and_depth", "content_only", "dip"])
def test_unmodified_reference_lightning_module_on_our_modules(case, monkeypatch, tmp_path):
    fake_engine.install(monkeypatch)
    ref = _load_reference_lightning_module()
    gold = torch.load(os.path.join(GOLDEN_DIR, f"{case}.pt"), weights_only=False)
    spec = gold["spec"]
    preset, sd, layers, view, style, hierarchical = build_inputs(spec)
    vgg_path = os.path.join(tmp_path, "vgg.pth")
    torch.save(sd, vgg_path)
    W, H = spec["tex_size"]
    mdl = ref.TextureOptimizationStyleTransferPipeline(
        W, H, hierarchical_texture=hierarchical, hierarchical_layers=len(layers), random_texture_init=True,
        style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
        use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
        style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
        angle_threshold=preset["angle_threshold"], learning_rate=spec["learning_rate"], decay_gamma=0.1,
        decay_step_size=30, loss_weights=dict(preset["loss_weights"]), tex_reg_weights=None, extra_args={})
    params = [m.data for m in mdl.texture.layers] if hierarchical else [mdl.texture.data]
    with torch.no_grad():
        for p, t in zip(params, layers):
            p.copy_(t)
    batch = view.as_batch()
    (opt,), _ = mdl.configure_optimizers()
    assert type(opt) is torch.optim.Adam                                     # the reference's optimizer, untouched

    # ---- teacher-forced: loss terms and dense gradients at the initial texture ----
    opt.zero_grad()
    res = mdl.training_step(batch, 0)
    res["loss"].backward()
    got = {k: float(mdl.loss_history[k]["train"][-1].reshape(-1)[0]) for k in ["content", "style", "tex_reg", "total"]}
    for k, v in gold["loss0"].items():
        assert rel(got[k], v) < 1e-4 or abs(got[k] - v) < 1e-6, (k, got[k], v)
    for l, (p, gg) in enumerate(zip(params, gold["grad0"])):
        assert (p.grad - gg).norm() <= 1e-4 * gg.norm() + 1e-12, (case, l)

    # ---- teacher-forced trajectory with the reference's own Adam ----
    if case == "dip":
        mdl.vgg_loss.gram_cache = {k: [] for k in mdl.vgg_loss.style_layers}      # as make_golden.py: restart the average
    (opt,), _ = mdl.configure_optimizers()
    for i in range(spec["steps"]):
        prev = gold["states"][i - 1] if i > 0 else None
        with torch.no_grad():
            for l, p in enumerate(params):
                p.copy_(prev["params"][l] if prev else layers[l])
                opt.state[p] = {"step": torch.tensor(float(i)),
                                "exp_avg": (prev["exp_avg"][l] if prev else torch.zeros_like(p)).clone(),
                                "exp_avg_sq": (prev["exp_avg_sq"][l] if prev else torch.zeros_like(p)).clone()}
        opt.zero_grad()
        mdl.training_step(batch, i)["loss"].backward()
        opt.step()
        got = {k: float(mdl.loss_history[k]["train"][-1].reshape(-1)[0]) for k in ["content", "style", "tex_reg", "total"]}
        for k, v in gold["traj"][i].items():
            assert rel(got[k], v) < 1e-3 or abs(got[k] - v) < 1e-6, (i, k, got[k], v)
        for l, p in enumerate(params):
            want = gold["states"][i]["params"][l]
            well = gold["states"][i]["exp_avg"][l].abs() > 1e-5 * gold["states"][i]["exp_avg"][l].abs().max()
            d = p.detach() - want
            assert d[well].norm() <= 1e-3 * want[well].norm(), (case, i, l)
