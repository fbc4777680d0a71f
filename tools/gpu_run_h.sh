#!/bin/bash
# same-box A/B of experiment knobs: each configuration twice, interleaved
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/h_probe.jsonl
for rep in 1 2; do
  for knob in 0 1; do
    SMB_CONV_IMPL=ph SMB_PH_KNOB=$knob PROBE_REPS=20 timeout 300 python tools/gpu_conv_probe.py 2>> gpurun_out/h_probe.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('knob $knob', d['igemm_fwd_ms'], {k:v['us'] for k,v in d['layers'].items()})" | tee -a gpurun_out/h_probe.jsonl
  done
  SMB_CONV_IMPL=tc PROBE_REPS=20 timeout 300 python tools/gpu_conv_probe.py 2>> gpurun_out/h_probe.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tc', d['igemm_fwd_ms'], {k:v['us'] for k,v in d['layers'].items()})" | tee -a gpurun_out/h_probe.jsonl
done
nvidia-smi --query-gpu=name,clocks.sm,power.draw,temperature.gpu --format=csv
