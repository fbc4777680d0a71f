"""Generate golden fixtures by running the REAL reference modules (imported from /root/reference, never copied).

    python tests/golden/make_golden.py            # writes tests/golden/<case>.pt

Only runs in the build container (needs /root/reference).  The fixtures pin the CPU oracle
(oracle/stylemesh_oracle.py) — see tests/test_oracle_vs_golden.py — and, through it, the CUDA path.

Shims needed to run the reference at all, none of which alters its arithmetic:
  1. `pytorch_lightning` is not installed -> a stub module (LightningModule = nn.Module + no-op logger).
  2. content_and_style_losses.py:298-299 builds its accumulators with torch.zeros(1, requires_grad=True).type_as(x);
     on CPU type_as is a no-op and `+=` on a leaf raises -> torch.zeros is wrapped to drop requires_grad during the
     call (on CUDA, where the reference was developed, type_as copies and the flag is lost the same way).
  3. VGG weights: seeded synthetic state_dict written to a temp file (no vgg_conv.pth offline).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest import mock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("STYLEMESH_REFERENCE", "/root/reference")


def _install_lightning_stub():
    pl = types.ModuleType("pytorch_lightning")

    class _Exp:
        def add_scalar(self, *a, **k):
            pass

        def add_scalars(self, *a, **k):
            pass

        def add_image(self, *a, **k):
            pass

    class _Logger:
        experiment = _Exp()

    class LightningModule(torch.nn.Module):
        current_epoch = 0
        logger = _Logger()

        def save_hyperparameters(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl


def _import_reference():
    _install_lightning_stub()
    sys.path.insert(0, REFERENCE)
    for name in [m for m in sys.modules if m == "model" or m.startswith("model.")]:
        del sys.modules[name]
    import model.model as ref_model                                   # noqa: E402
    import model.losses.content_and_style_losses as ref_cs            # noqa: E402
    assert os.path.abspath(ref_model.__file__).startswith(os.path.abspath(REFERENCE)), ref_model.__file__
    return ref_model, ref_cs


def golden_case_specs():
    """Shared with tests/test_oracle_vs_golden.py: how each case's inputs are (re)generated from seeds."""
    base = dict(rgb_size=(48, 64), tex_size=(64, 64), style_size=(96, 80), vgg_seed=0, vgg_bias_scale=0.05,
                tex_seed=11, view_seed=1000, style_seed=7, steps=3, learning_rate=1.0)
    cases = {}
    for preset in ["only2D", "with_angle", "with_angle_and_depth", "dip", "content_only"]:
        c = dict(base)
        c["preset"] = preset
        cases[preset] = c
    cases["with_angle_and_depth"]["level_sizes"] = [(48, 64), (72, 96), (96, 128)]
    return cases


def build_inputs(spec):
    sys.path.insert(0, REPO) if REPO not in sys.path else None
    from stylemesh_b200 import synthetic as syn
    preset = dict(syn.PRESETS[spec["preset"]])
    level_sizes = spec.get("level_sizes", [spec["rgb_size"]])
    n_layers = min(preset["hierarchical_layers"], 3)
    sd = syn.make_vgg_state_dict(spec["vgg_seed"], bias_scale=spec["vgg_bias_scale"])
    layers = syn.make_texture_layers(spec["tex_seed"], spec["tex_size"][0], spec["tex_size"][1], n_layers)
    view = syn.make_view(spec["view_seed"], spec["rgb_size"], level_sizes)
    style = syn.make_style_image(spec["style_seed"], *spec["style_size"])
    hierarchical = spec["preset"] != "content_only"
    return preset, sd, layers, view, style, hierarchical


def run_reference(spec):
    ref_model, ref_cs = _import_reference()
    preset, sd, layers, view, style, hierarchical = build_inputs(spec)
    real_zeros = torch.zeros

    def zeros_no_grad(*a, requires_grad=False, **k):
        return real_zeros(*a, **k)

    with tempfile.TemporaryDirectory() as td:
        vgg_path = os.path.join(td, "vgg_synth.pth")
        torch.save(sd, vgg_path)
        W, H = spec["tex_size"]
        mdl = ref_model.TextureOptimizationStyleTransferPipeline(
            W, H, hierarchical_texture=hierarchical, hierarchical_layers=len(layers), random_texture_init=True,
            style_image=style.clone(), style_weights=list(preset["style_weights"]),
            vgg_gatys_model_path=vgg_path, use_angle_weight=preset["use_angle_weight"],
            use_depth_scaling=preset["use_depth_scaling"], style_pyramid_mode=preset["style_pyramid_mode"],
            gram_mode=preset["gram_mode"], angle_threshold=preset["angle_threshold"],
            learning_rate=spec["learning_rate"], decay_gamma=0.1, decay_step_size=30,
            loss_weights=dict(preset["loss_weights"]), tex_reg_weights=None, extra_args={})
    with torch.no_grad():
        if hierarchical:
            for mod, t in zip(mdl.texture.layers, layers):
                mod.data.copy_(t)
        else:
            mdl.texture.data.copy_(layers[0])
    params = [m.data for m in mdl.texture.layers] if hierarchical else [mdl.texture.data]
    batch = view.as_batch()
    (opt,), _ = mdl.configure_optimizers()

    out = {"spec": spec}
    with mock.patch.object(ref_cs.torch, "zeros", zeros_no_grad):
        # --- teacher-forced: losses and dense gradient at the initial texture ---
        opt.zero_grad()
        res = mdl.training_step(batch, 0)
        res["loss"].backward()
        out["loss0"] = {k: float(mdl.loss_history[k]["train"][-1].reshape(-1)[0]) for k in
                        ["content", "style", "tex_reg", "total"]}
        out["grad0"] = [p.grad.detach().clone() for p in params]
        out["style_target_r11_l0"] = mdl.vgg_loss.style_targets[0][0].detach().clone()
        out["style_target_sums"] = [[float(mdl.vgg_loss.style_targets[i][l].sum()) for l in range(5)]
                                    for i in range(len(mdl.vgg_loss.style_layers))]
        if spec["preset"] == "dip":
            mdl.vgg_loss.gram_cache = {k: [] for k in mdl.vgg_loss.style_layers}   # restart the running average
        # --- free-running: `steps` Adam steps from the same initial texture ---
        with torch.no_grad():
            for p, t in zip(params, layers):
                p.copy_(t)
        (opt,), _ = mdl.configure_optimizers()
        traj, states = [], []
        for i in range(spec["steps"]):
            opt.zero_grad()
            res = mdl.training_step(batch, i)
            res["loss"].backward()
            opt.step()
            traj.append({k: float(mdl.loss_history[k]["train"][-1].reshape(-1)[0]) for k in
                         ["content", "style", "tex_reg", "total"]})
            # optimiser state after step i: lets a test replay step i+1 teacher-forced from the reference's state
            states.append({"params": [p.detach().clone() for p in params],
                           "exp_avg": [opt.state[p]["exp_avg"].clone() for p in params],
                           "exp_avg_sq": [opt.state[p]["exp_avg_sq"].clone() for p in params]})
        out["traj"] = traj
        out["states"] = states
        out["final_layers"] = [p.detach().clone() for p in params]
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, spec in golden_case_specs().items():
        res = run_reference(spec)
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(res, path)
        print(f"[golden] {name}: loss0={res['loss0']} -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
