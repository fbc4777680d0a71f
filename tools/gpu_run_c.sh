#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/c_trace.jsonl
for layer in 5 4 12 1; do
  PROBE_LAYER=$layer timeout 300 python tools/gpu_trace_probe.py >> gpurun_out/c_trace.jsonl 2>> gpurun_out/c_trace.err
done
PROBE_LAYER=5 SMB_IGEMM_DEBUG=4 timeout 300 python tools/gpu_trace_probe.py >> gpurun_out/c_trace.jsonl 2>> gpurun_out/c_trace.err
PROBE_LAYER=5 SMB_IGEMM_DEBUG=2 timeout 300 python tools/gpu_trace_probe.py >> gpurun_out/c_trace.jsonl 2>> gpurun_out/c_trace.err
python - <<'PY'
import json
for line in open("gpurun_out/c_trace.jsonl"):
    d = json.loads(line)
    print(d["layer"], d["dbg"], d["ctas"])
    for k, v in d["summary"].items():
        print("   %-14s min %10.0f med %10.0f max %10.0f" % (k, v["min"], v["med"], v["max"]))
PY
tail -5 gpurun_out/c_trace.err
