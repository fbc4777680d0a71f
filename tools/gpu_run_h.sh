#!/bin/bash
# same-box A/B of experiment knobs
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_vgg_units.py -q -m gpu -x -k "ph" --timeout 300 -p no:cacheprovider > gpurun_out/h_units.log 2>&1
echo "units exit $?"; tail -n 3 gpurun_out/h_units.log
: > gpurun_out/h_probe.jsonl
for cfg in "SMB_PH_NO_RESIDENT=0" "SMB_PH_NO_RESIDENT=1" "SMB_PH_NO_RESIDENT=0" "SMB_PH_KNOB=1"; do
    env SMB_CONV_IMPL=ph $cfg PROBE_REPS=20 timeout 300 python tools/gpu_conv_probe.py 2>> gpurun_out/h_probe.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', d['igemm_fwd_ms'], {k:v['us'] for k,v in d['layers'].items()})" | tee -a gpurun_out/h_probe.jsonl
done
