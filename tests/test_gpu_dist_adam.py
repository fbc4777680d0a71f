"""Multi-GPU: the fused reduce-scatter + Adam + all-gather kernel (smb_dist_adam_step) against NCCL all_reduce + the
single-GPU Adam kernel, 2 ranks on one box (skipped on single-GPU boxes; the CPU suite covers the host logic of the
view-sharded step with gloo, tests/test_ddp_gloo_cpu.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_dist_adam_equals_allreduce_plus_adam():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(REPO, "tools", "dist_adam_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["ok"] and out["max_abs_diff_vs_allreduce_path"] == 0.0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("preset,view", [("only2D", "120x160"), ("with_angle_and_depth", "96x128")])
def test_two_rank_pipeline_matches_mean_gradient_oracle(preset, view):
    """the whole step on 2 real ranks (own view each, fused exchange) vs OraclePipeline.step_views, teacher-forced"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29543", os.path.join(REPO, "tools", "dist_pipeline_check.py"),
           "--preset", preset, "--view", view, "--texture", "512", "--steps", "3"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines and lines[-1]["ok"], res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cli_on_two_ranks_writes_one_log_version_and_a_checkpoint(tmp_path):
    """`torchrun --nproc-per-node 2 -m model.optimize ...` (the reference's CLI, one process per GPU): views are sharded
    over the ranks, rank 0 alone picks the log version, logs, validates, exports the texture and writes the checkpoint
    (gathering the sharded Adam moments); the run can be resumed from it on ONE GPU."""
    argv = ["--gpus", "2", "--dataset", "synthetic", "--resize_size", "96", "--texture_size", "256,256", "--max_images", "6",
            "--hierarchical", "--hierarchical_layers", "3", "--loss_weight", "content=7e1", "--loss_weight", "style=1e-4",
            "--style_weights=1000,1000,10,10,1000", "--loss_weight", "tex_reg=5e3", "--vgg_gatys_model_path", "synthetic:0",
            "--learning_rate", "1", "--max_epochs", "2", "--train_split", "0.67", "--val_split", "0.33", "--sampler_mode",
            "repeat", "--index_repeat", "2", "--save_texture", "--split_mode", "sequential", "--style_image_path",
            "synthetic:96:80", "--default_root_dir", str(tmp_path), "--random_texture_init", "--style_pyramid_mode", "single",
            "--gram_mode", "current", "--angle_threshold", "3000", "--pyramid_levels", "1", "--no_depth_scaling",
            "--no_angle_weight"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29547", "-m", "model.optimize", *argv]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    import glob
    versions = glob.glob(os.path.join(str(tmp_path), "lightning_logs", "version_*"))
    assert len(versions) == 1, versions
    rows = [json.loads(l) for l in open(os.path.join(versions[0], "scalars.jsonl"))]
    tot = [r["value"] for r in rows if r["tag"] == "Batch/Loss/train/total"]
    assert len(tot) == 2 * (4 * 2 // 2)                      # 2 epochs x (4 train views x repeat 2) / 2 ranks, rank 0's share
    ckpts = glob.glob(os.path.join(versions[0], "checkpoints", "*.ckpt"))
    assert len(ckpts) == 1
    ck = torch.load(ckpts[0], map_location="cpu", weights_only=False)
    st = ck["optimizer_states"][0]["state"]
    assert ck["epoch"] == 2 and int(st[0]["step"]) == 8
    for i in range(3):                                        # every slice of the sharded moments was gathered
        v = st[i]["exp_avg_sq"].reshape(-1)
        n = v.numel()
        assert float(v[: n // 2].abs().sum()) > 0 and float(v[n // 2:].abs().sum()) > 0, i
    assert glob.glob(os.path.join(versions[0], "*texture.jpg"))
    # resume on a single GPU from the 2-rank checkpoint
    from model.optimize import build_parser, main
    one = [a for a in argv]
    one[one.index("--max_epochs") + 1] = "3"
    one[one.index("--gpus") + 1] = "1"
    res1 = main(build_parser().parse_args(one + ["--resume_from_checkpoint", ckpts[0]]))
    assert res1.trainer.start_epoch == 2 and res1.trainer.optimizers[0]._steps == 8 + 8
