#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_vgg_units.py -q -m gpu -x -k "ph" --timeout 300 -p no:cacheprovider > gpurun_out/k_units.log 2>&1
echo "units exit $?"; tail -n 3 gpurun_out/k_units.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize_properties.py -q -m gpu -k "not simt" --timeout 600 -p no:cacheprovider > gpurun_out/k_pipe.log 2>&1
echo "pipeline exit $?"; tail -n 5 gpurun_out/k_pipe.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/k_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d["e2e"]["value"])
PY
