#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_vgg_units.py -q -m gpu -k "ph" -x --timeout 300 -p no:cacheprovider > gpurun_out/d_units.log 2>&1
echo "units exit $?"; tail -n 25 gpurun_out/d_units.log
: > gpurun_out/d_probe.jsonl
SMB_CONV_IMPL=ph timeout 300 python tools/gpu_conv_probe.py >> gpurun_out/d_probe.jsonl 2>> gpurun_out/d_probe.err
SMB_CONV_IMPL=ph SMB_PH_DIRECT_STORES=1 timeout 300 python tools/gpu_conv_probe.py >> gpurun_out/d_probe.jsonl 2>> gpurun_out/d_probe.err
cat gpurun_out/d_probe.jsonl; tail -5 gpurun_out/d_probe.err
