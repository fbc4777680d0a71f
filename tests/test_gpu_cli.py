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
                 "--pyramid_levels", "1", "--no_depth_scaling", "--no_angle_weight"]
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
