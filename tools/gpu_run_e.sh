#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/e_trace.jsonl
for layer in 5 12 1 3; do
  SMB_CONV_IMPL=ph PROBE_LAYER=$layer timeout 300 python tools/gpu_trace_probe.py >> gpurun_out/e_trace.jsonl 2>> gpurun_out/e_trace.err
done
python - <<'PY'
import json
for line in open("gpurun_out/e_trace.jsonl"):
    d = json.loads(line)
    print(d["layer"], d["impl"], d["ctas"])
    for k, v in d["summary"].items():
        print("   %-14s min %10.0f med %10.0f max %10.0f" % (k, v["min"], v["med"], v["max"]))
PY
tail -5 gpurun_out/e_trace.err
