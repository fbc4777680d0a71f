#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/s_trace.jsonl
for layer in 1 3; do
  SMB_CONV_IMPL=ph PROBE_LAYER=$layer timeout 300 python tools/gpu_trace_probe.py >> gpurun_out/s_trace.jsonl 2>> gpurun_out/s_trace.err
done
for knob in 2 4 8; do
  SMB_PH_KNOB=$knob SMB_CONV_IMPL=ph PROBE_LAYER=1 timeout 300 python tools/gpu_trace_probe.py >> gpurun_out/s_trace.jsonl 2>> gpurun_out/s_trace.err
done
python - <<'PY'
import json
for line in open("gpurun_out/s_trace.jsonl"):
    d = json.loads(line)
    print(d["layer"], d["impl"], d["ctas"])
    for k, v in d["summary"].items():
        print("   %-14s min %10.0f med %10.0f max %10.0f" % (k, v["min"], v["med"], v["max"]))
PY
