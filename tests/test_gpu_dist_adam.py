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
