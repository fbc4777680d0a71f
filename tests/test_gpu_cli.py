"""The reference's command line (`python -m model.optimize ...`, scripts/train/optimize_texture_*.sh) end to end on
the B200 path with the in-memory synthetic scene: flags parse, the Trainer loop runs train + val epochs, StepLR
steps, textures are exported, and the optimisation actually reduces the loss."""
import glob
import json
import os

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("family", ["only2D", "with_angle_and_depth"])
def test_cli_runs_and_loss_decreases(family, tmp_path):
    from model.optimize import build_parser, main
    argv = ["--gpus", "1", "--dataset", "synthetic", "--resize_size", "96", "--texture_size", "256,256",
            "--max_images", "4", "--hierarchical", "--hierarchical_layers", "3",
            "--loss_weight", "content=7e1", "--loss_weight", "style=1e-4", "--style_weights=1000,1000,10,10,1000",
            "--loss_weight", "tex_reg=5e3", "--vgg_gatys_model_path", "synthetic:0", "--learning_rate", "1",
            "--decay_step_size", "3", "--max_epochs", "3", "--train_split", "0.75", "--val_split", "0.25",
            "--sampler_mode", "repeat", "--index_repeat", "4", "--save_texture", "--split_mode", "sequential",
            "--style_image_path", "synthetic:96:80", "--default_root_dir", str(tmp_path), "--random_texture_init"]
    if family == "only2D":
        argv += ["--style_pyramid_mode", "single", "--gram_mode", "current", "--angle_threshold", "3000",
                 "--pyramid_levels", "1", "--no_depth_scaling", "--no_angle_weight", "--renderer_mipmap", "cuda",
                 "--preview_views", "2"]
    else:
        argv += ["--style_pyramid_mode", "multi", "--gram_mode", "current", "--angle_threshold", "30",
                 "--pyramid_levels", "3"]
    model = main(build_parser().parse_args(argv))
    logs = glob.glob(os.path.join(str(tmp_path), "lightning_logs", "version_*"))
    assert len(logs) == 1
    assert glob.glob(os.path.join(logs[0], "*texture.jpg")), "epoch-end texture export missing"
    rows = [json.loads(l) for l in open(os.path.join(logs[0], "scalars.jsonl"))]
    tot = [r["value"] for r in rows if r["tag"] == "Batch/Loss/train/total"]
    assert len(tot) == 3 * 3 * 4                      # 3 epochs x 3 train views x index_repeat 4
    assert tot[-1] < 0.9 * tot[0], (tot[0], tot[-1])
    assert any(r["tag"] == "Batch/Loss/val/total" for r in rows)
    previews = glob.glob(os.path.join(logs[0], "preview", "*.jpg"))
    assert len(previews) == (2 if family == "only2D" else 0)      # headless styled views (--renderer_mipmap cuda)
    if previews:
        from PIL import Image
        import numpy as np
        img = np.asarray(Image.open(previews[0]))
        assert img.shape == (96, 128, 3) and img.std() > 1.0


def _argv(tmp, epochs, extra=()):
    return ["--gpus", "1", "--dataset", "synthetic", "--resize_size", "96", "--texture_size", "256,256",
            "--max_images", "4", "--hierarchical", "--hierarchical_layers", "3",
            "--loss_weight", "content=7e1", "--loss_weight", "style=1e-4", "--style_weights=1000,1000,10,10,1000",
            "--loss_weight", "tex_reg=5e3", "--vgg_gatys_model_path", "synthetic:0", "--learning_rate", "1",
            "--decay_step_size", "1", "--decay_gamma", "0.5", "--max_epochs", str(epochs), "--train_split", "0.75",
            "--val_split", "0.25", "--sampler_mode", "repeat", "--index_repeat", "3", "--split_mode", "sequential",
            "--style_image_path", "synthetic:96:80", "--default_root_dir", str(tmp), "--random_texture_init",
            "--style_pyramid_mode", "single", "--gram_mode", "current", "--angle_threshold", "3000",
            "--pyramid_levels", "1", "--no_depth_scaling", "--no_angle_weight", *extra]


def _train_totals(log_dir):
    rows = [json.loads(l) for l in open(os.path.join(log_dir, "scalars.jsonl"))]
    return [r["value"] for r in rows if r["tag"] == "Batch/Loss/train/total"]


def test_checkpoint_every_epoch_and_resume(tmp_path):
    """Lightning's implicit ModelCheckpoint + --resume_from_checkpoint (reference: model/model.py:69-72,
    optimize.py:30,241): a run interrupted after epoch 1 and resumed continues like the uninterrupted run - same
    epoch counter, StepLR state, Adam moments and bias-correction step."""
    import torch
    from model.optimize import build_parser, main
    torch.manual_seed(0)
    full = main(build_parser().parse_args(_argv(tmp_path / "full", 3)))
    torch.manual_seed(0)
    part = main(build_parser().parse_args(_argv(tmp_path / "part", 2)))
    ckpts = glob.glob(os.path.join(str(tmp_path / "part"), "lightning_logs", "version_0", "checkpoints", "*.ckpt"))
    assert len(ckpts) == 1 and "epoch=1" in os.path.basename(ckpts[0]), ckpts          # latest only, like Lightning
    ck = torch.load(ckpts[0], map_location="cpu", weights_only=False)
    assert ck["epoch"] == 2 and ck["global_step"] == 2 * 3 * 3
    st = ck["optimizer_states"][0]["state"]
    assert len(st) == 3 and int(st[0]["step"]) == 18 and float(st[0]["exp_avg_sq"].abs().sum()) > 0
    for k, v in ck["state_dict"].items():                                              # texels as trained
        assert torch.equal(v, part.state_dict()[k].cpu()), k

    torch.manual_seed(123)                      # a different random init: everything must come from the checkpoint
    res = main(build_parser().parse_args(_argv(tmp_path / "part", 3, ["--resume_from_checkpoint", ckpts[0]])))
    assert res.trainer.start_epoch == 2 and res.trainer.global_step == 27
    resumed = _train_totals(os.path.join(str(tmp_path / "part"), "lightning_logs", "version_1"))
    whole = _train_totals(os.path.join(str(tmp_path / "full"), "lightning_logs", "version_0"))
    assert len(resumed) == 9 and len(whole) == 27
    # free-running fp32 trajectories of two runs drift apart (atomics order); the loss stays close
    for a, b in zip(resumed, whole[18:]):
        assert abs(a - b) <= 5e-3 * abs(b), (resumed, whole[18:])
    (opt_lr,) = [g["lr"] for g in res.trainer.optimizers[0].param_groups]
    assert abs(opt_lr - 0.125) < 1e-12          # StepLR(gamma 0.5, every epoch) after 3 epochs, not 2 restarts from 1.0


def test_resume_from_missing_checkpoint_raises(tmp_path):
    from model.optimize import build_parser, main
    with pytest.raises(FileNotFoundError):
        main(build_parser().parse_args(_argv(tmp_path, 1, ["--resume_from_checkpoint", str(tmp_path / "nope.ckpt")])))
