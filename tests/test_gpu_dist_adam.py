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
