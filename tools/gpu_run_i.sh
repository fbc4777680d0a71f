#!/bin/bash
bash tools/gpu_check.sh
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/i_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step", "gpu_launches")}, d["e2e"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
PY
