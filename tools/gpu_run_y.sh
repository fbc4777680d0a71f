#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)|^E  |watchdog" gpurun_out/$name.log | head -n 30; }
run y_units tests/test_gpu_vgg_units.py -k "gram"
run y_pipe tests/test_gpu_pipeline.py -k "not simt"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/y_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d["e2e"]["value"], d["with_cached_content_targets"]["value"])
PY
