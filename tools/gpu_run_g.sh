#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_vgg_units.py -q -m gpu -x --timeout 300 -p no:cacheprovider > gpurun_out/g_units.log 2>&1
echo "units exit $?"; tail -n 5 gpurun_out/g_units.log
: > gpurun_out/g_trace.jsonl
for layer in 5 12; do
  SMB_CONV_IMPL=ph PROBE_LAYER=$layer timeout 300 python tools/gpu_trace_probe.py >> gpurun_out/g_trace.jsonl 2>> gpurun_out/g_trace.err
done
python - <<'PY'
import json
for line in open("gpurun_out/g_trace.jsonl"):
    d = json.loads(line)
    print(d["layer"], d["impl"], d["ctas"])
    for k, v in d["summary"].items():
        print("   %-14s min %10.0f med %10.0f max %10.0f" % (k, v["min"], v["med"], v["max"]))
PY
SMB_CONV_IMPL=ph timeout 300 python tools/gpu_conv_probe.py 2>> gpurun_out/g_probe.err | tee gpurun_out/g_probe.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-impl ph > gpurun_out/g_bench_ph.json 2> gpurun_out/g_bench_ph.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/g_bench_ph.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d["e2e"]["value"])
PY
