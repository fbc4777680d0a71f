#!/bin/bash
# round-2 GPU call E: suite with the K5 kernels + pooled side output; A/B of the side output; host cost per step
set -u
mkdir -p gpurun_out
export SMB_PARITY_LOG=gpurun_out/r02e_parity_stats.jsonl
rm -f $SMB_PARITY_LOG
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02e_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest_gpu.log
tail -15 gpurun_out/r02e_pytest_gpu.log
unset SMB_PARITY_LOG
SMB_PH_POOL_SIDE=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity --sustained-s 0 > gpurun_out/r02e_bench_c2_poolside0.json 2>/dev/null
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity --sustained-s 0 > gpurun_out/r02e_bench_c2_poolside1.json 2>/dev/null
python - <<'PY'
import json
for n in ["poolside0", "poolside1"]:
    try:
        d = json.load(open(f"gpurun_out/r02e_bench_c2_{n}.json"))
        print(n, round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "ms; host", round(d["host_enqueue_ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["kernel_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 300 python tools/host_profile.py > gpurun_out/r02e_host_profile_c2.txt 2>/dev/null; head -3 gpurun_out/r02e_host_profile_c2.txt
timeout 300 python tools/host_profile.py --preset with_angle_and_depth --view 256x341 > gpurun_out/r02e_host_profile_c3.txt 2>/dev/null; head -22 gpurun_out/r02e_host_profile_c3.txt
