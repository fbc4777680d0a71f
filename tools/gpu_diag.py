"""Numerical diagnostics on the GPU box (not a test): where does the CUDA path's error come from, and how does it
propagate to losses / gradients / texels?  Prints a JSON report to gpurun_out/diag.json.

    python tools/gpu_diag.py
"""
import json
import os
import sys
import tempfile

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))

from make_golden import build_inputs, golden_case_specs  # noqa: E402
from oracle import stylemesh_oracle as orc  # noqa: E402

report = {}


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def conv_units():
    from stylemesh_b200 import engine as eng
    out = []
    for (cin, cout, h, w) in [(64, 64, 60, 80), (256, 256, 30, 40), (512, 512, 15, 20), (512, 512, 30, 40)]:
        for positive in (False, True):
            g = torch.Generator().manual_seed(cin + h)
            x = torch.randn(cin, h, w, generator=g) * 50
            wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
            if positive:
                x, wt = x.abs(), wt.abs()
            ref = F.conv2d(x.double().unsqueeze(0), wt.double(), None, padding=1)[0]
            ref32 = F.conv2d(x.unsqueeze(0), wt, None, padding=1)[0]
            row = {"shape": [cin, cout, h, w], "positive": positive, "cpu_fp32_rel": rel(ref32, ref),
                   "cpu_fp32_bias": float((ref32.double() - ref).sum() / ref.abs().sum())}
            for impl, name in ((0, "simt"), (1, "tc")):
                y = eng.unit_conv3x3(impl, x.cuda(), wt, None, relu=False).cpu()
                row[name + "_rel"] = rel(y, ref)
                row[name + "_bias"] = float((y.double() - ref).sum() / ref.abs().sum())
            out.append(row)
            print(row, flush=True)
    report["conv_units"] = out


def pipeline(case="only2D"):
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    spec = golden_case_specs()[case]
    gold = torch.load(os.path.join(REPO, "tests", "golden", f"{case}.pt"), weights_only=False)
    preset, sd, layers, view, style, hierarchical = build_inputs(spec)
    keys = ["r11", "r12", "r21", "r22", "r31", "r32", "r33", "r34", "r41", "r42", "r43", "r44", "r51"]
    res = {}
    # oracle reference pieces (fp64 features for an exact yardstick)
    pred_ref = orc.texture_sample([l.clone() for l in layers], view.uvs[0])
    sd64 = {k: v.double() for k, v in sd.items()}
    f64 = orc.vgg_forward(sd64, pred_ref.double(), keys, as_written=False)
    f32 = orc.vgg_forward(sd, pred_ref, keys, as_written=False)
    res["cpu_fp32_feature_rel"] = {k: rel(f32[k], f64[k]) for k in keys}
    res["cpu_fp32_mask_mismatch"] = {k: float(((f32[k] > 0) != (f64[k] > 0)).float().mean()) for k in keys}
    for impl in ("simt", "tc"):
        os.environ["SMB_CONV_IMPL"] = impl
        os.environ["SMB_GRAM_IMPL"] = impl
        with tempfile.TemporaryDirectory() as td:
            vgg_path = os.path.join(td, "vgg.pth")
            torch.save(sd, vgg_path)
            W, H = spec["tex_size"]
            mdl = TextureOptimizationStyleTransferPipeline(
                W, H, hierarchical_texture=hierarchical, hierarchical_layers=len(layers), random_texture_init=True,
                style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
                use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
                style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
                angle_threshold=preset["angle_threshold"], learning_rate=1.0, loss_weights=dict(preset["loss_weights"]),
                save_texture=False)
        mdl.cuda()
        mods = list(mdl.texture.layers)
        with torch.no_grad():
            for m, t in zip(mods, layers):
                m.data.copy_(t.cuda())
        batch = view.to("cuda").as_batch()
        r = {}
        feats = mdl.vgg_loss.vgg(pred_ref.cuda(), keys)
        r["feature_rel_vs_fp64"] = {k: rel(feats[k].cpu(), f64[k]) for k in keys}
        r["feature_bias_vs_fp64"] = {k: float((feats[k].cpu().double() - f64[k]).sum() / f64[k].abs().sum()) for k in keys}
        r["mask_mismatch_vs_fp64"] = {k: float(((feats[k].cpu() > 0) != (f64[k] > 0)).float().mean()) for k in keys}
        mdl.training_step(batch, 0)
        buf = mdl._loss_buf.cpu()
        got = {"style": float(buf[0]), "content": float(buf[1]), "tex_reg": float(buf[2]), "total": float(buf[3])}
        r["loss_rel"] = {k: abs(got[k] - v) / max(abs(v), 1e-12) for k, v in gold["loss0"].items()}
        lam = float(mdl.loss_weights.get("tex_reg", 0.0))
        gr = []
        for l, (g, gg) in enumerate(zip(mdl._grad_tensors(), gold["grad0"])):
            x = mods[l].data.detach().cpu().clamp(orc.CLAMP_LO, orc.CLAMP_HI)
            reg = lam * mdl.tex_reg_weights[l] * 2.0 * x / x.numel()
            gr.append(rel(g.cpu() + reg, gg))
        r["grad_rel"] = gr
        # one Adam step from the same state: texel agreement
        (opt,), _ = mdl.configure_optimizers()
        opt.step()
        tex = []
        for l, m in enumerate(mods):
            want = gold["states"][0]["params"][l]
            d = (m.data.detach().cpu() - want)
            g0 = gold["grad0"][l]
            well = g0.abs() > 1e-2 * g0.abs().mean()
            tex.append({"rel_all": float(d.norm() / want.norm()), "rel_well": float(d[well].norm() / want[well].norm()),
                        "frac_texels_off_by_gt_0.5": float((d.abs() > 0.5).float().mean()),
                        "frac_well": float(well.float().mean())})
        r["texels_after_step1"] = tex
        res[impl] = r
        print(impl, json.dumps(r)[:3000], flush=True)
        # free-running loss curve vs the oracle (10 steps)
        if impl == "tc":
            with torch.no_grad():
                for m, t in zip(mods, layers):
                    m.data.copy_(t.cuda())
            st = mdl._ensure_fused_state()
            st["exp_avg"].zero_(); st["exp_avg_sq"].zero_(); st["grad"].zero_()
            (opt,), _ = mdl.configure_optimizers()
            loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                          angle_threshold=preset["angle_threshold"],
                                          style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
                                          as_written=False)
            loss.set_style_image(style.unsqueeze(0))
            cfg = orc.OracleConfig(use_angle_weight=preset["use_angle_weight"],
                                   use_depth_scaling=preset["use_depth_scaling"],
                                   loss_weights=dict(preset["loss_weights"]), hierarchical=hierarchical, learning_rate=1.0)
            pipe = orc.OraclePipeline(layers, loss, cfg)
            curve = []
            cb = view.as_batch()
            for i in range(12):
                mdl.training_step(batch, i)
                opt.step()
                ours = float(mdl._loss_buf[3])
                ref = pipe.step(cb)["total"]
                curve.append({"step": i, "ours": ours, "ref": ref, "rel": abs(ours - ref) / abs(ref)})
            res["free_running_loss_curve"] = curve
            print(json.dumps(curve), flush=True)
    report["pipeline_" + case] = res


def cpu_threads():
    import time
    info = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    for f in ("/sys/fs/cgroup/cpu.max", "/proc/loadavg"):
        try:
            info[f] = open(f).read().strip()
        except Exception as e:
            info[f] = str(e)
    x = torch.randn(1, 256, 120, 160)
    w = torch.randn(256, 256, 3, 3)
    for nt in (8, 16, 32, 64, 128):
        if nt > (os.cpu_count() or 1):
            continue
        torch.set_num_threads(nt)
        F.conv2d(x, w, padding=1)
        t = time.perf_counter()
        for _ in range(3):
            F.conv2d(x, w, padding=1)
        dt = (time.perf_counter() - t) / 3
        info[f"conv_gflops_{nt}t"] = 2 * 9 * 256 * 256 * 120 * 160 / dt / 1e9
    report["cpu"] = info
    print(info, flush=True)


if __name__ == "__main__":
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    cpu_threads()
    conv_units()
    pipeline("only2D")
    json.dump(report, open(os.path.join(REPO, "gpurun_out", "diag.json"), "w"), indent=1)
