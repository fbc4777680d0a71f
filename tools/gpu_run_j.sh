#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_cli.py -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/j_cli.log 2>&1
echo "cli exit $?"; tail -n 3 gpurun_out/j_cli.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
echo "bench exit $?"; tail -3 gpurun_out/j_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/j_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["with_cached_content_targets"])
PY
