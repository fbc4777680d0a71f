"""bench.py contract on a box without a GPU: the reference arm (the oracle port on the host cores) prints exactly one
JSON line with the keys the driver reads, and our arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--texture", "128", "--layers", "2",
                "--view", "64x96", "--style", "96x80", "--views-per-gpu", "2"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_our_arm_has_no_cpu_path():
    if torch.cuda.is_available():
        return
    res = _run(["--steps", "1", "--warmup", "1"])
    assert res.returncode != 0 and "no CPU path" in (res.stderr + res.stdout)
