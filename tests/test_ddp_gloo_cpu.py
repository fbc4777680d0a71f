"""View-sharded data parallelism: two gloo ranks on CPU, each with its own view, one all-reduce of the
flat texture gradient per step inside FusedTextureAdam.step, must equal the single-process oracle that averages the
per-view gradients (OraclePipeline.step_views).  The engine is the CPU emulation of tests/fake_engine.py; the NCCL
path on real GPUs is exercised by bench.py --gpus N."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Patch:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _build(case):
    for p in (REPO, HERE, os.path.join(HERE, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from make_golden import build_inputs, golden_case_specs
    spec = golden_case_specs()[case]
    return spec, build_inputs(spec)


def _worker(rank, world, port, tmpdir, case, steps):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec, (preset, sd, layers, view, style, hierarchical) = _build(case)
    import fake_engine
    fake_engine.install(_Patch())
    from stylemesh_b200 import synthetic as syn
    from stylemesh_b200.model.model import TextureOptimizationStyleTransferPipeline
    vgg_path = os.path.join(tmpdir, f"vgg_{rank}.pth")
    torch.save(sd, vgg_path)
    W, H = spec["tex_size"]
    mdl = TextureOptimizationStyleTransferPipeline(
        W, H, hierarchical_texture=True, hierarchical_layers=len(layers), random_texture_init=True,
        style_image=style.clone(), style_weights=list(preset["style_weights"]), vgg_gatys_model_path=vgg_path,
        use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
        style_pyramid_mode=preset["style_pyramid_mode"], gram_mode=preset["gram_mode"],
        angle_threshold=preset["angle_threshold"], learning_rate=1.0, loss_weights=dict(preset["loss_weights"]),
        save_texture=False)
    with torch.no_grad():
        for m, t in zip(mdl.texture.layers, layers):
            m.data.copy_(t)
    my_view = syn.make_view(spec["view_seed"] + rank, spec["rgb_size"], spec.get("level_sizes", [spec["rgb_size"]]))
    (opt,), _ = mdl.configure_optimizers()
    for i in range(steps):
        opt.zero_grad()
        mdl.training_step(my_view.as_batch(), i)["loss"].backward()
        opt.step()
    torch.save([m.data.detach().clone() for m in mdl.texture.layers], os.path.join(tmpdir, f"tex_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["only2D"])
def test_two_rank_view_sharding_matches_mean_gradient_oracle(case, tmp_path):
    world, steps = 2, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), case, steps), nprocs=world, join=True)
    t0 = torch.load(os.path.join(tmp_path, "tex_0.pt"))
    t1 = torch.load(os.path.join(tmp_path, "tex_1.pt"))
    for a, b in zip(t0, t1):
        assert torch.equal(a, b), "replicas must stay bit-identical after the all-reduce"

    spec, (preset, sd, layers, view, style, hierarchical) = _build(case)
    from oracle import stylemesh_oracle as orc
    from stylemesh_b200 import synthetic as syn
    loss = orc.StyleContentOracle(vgg_params=sd, style_weights=list(preset["style_weights"]),
                                  angle_threshold=preset["angle_threshold"],
                                  style_pyramid_mode=preset["style_pyramid_mode"], as_written=False)
    loss.set_style_image(style.unsqueeze(0))
    cfg = orc.OracleConfig(use_angle_weight=preset["use_angle_weight"], use_depth_scaling=preset["use_depth_scaling"],
                           loss_weights=dict(preset["loss_weights"]), hierarchical=True, learning_rate=1.0)
    pipe = orc.OraclePipeline(layers, loss, cfg)
    views = [syn.make_view(spec["view_seed"] + r, spec["rgb_size"], spec.get("level_sizes", [spec["rgb_size"]]))
             for r in range(world)]
    for _ in range(steps):
        pipe.step_views([v.as_batch() for v in views])
    for a, b in zip(t0, pipe.layers):
        assert (a - b.detach()).norm() <= 2e-3 * b.detach().norm(), float((a - b.detach()).norm() / b.detach().norm())
