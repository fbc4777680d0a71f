#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/b_probe.jsonl
for dbg in 0 1 2 3 4; do
  SMB_CONV_IMPL=tc SMB_IGEMM_DEBUG=$dbg timeout 300 python tools/gpu_conv_probe.py >> gpurun_out/b_probe.jsonl 2>> gpurun_out/b_probe.err
done
SMB_CONV_IMPL=tc SMB_IGEMM_MAX_BN=256 timeout 300 python tools/gpu_conv_probe.py >> gpurun_out/b_probe.jsonl 2>> gpurun_out/b_probe.err
SMB_CONV_IMPL=tc SMB_IGEMM_MAX_BN=64 timeout 300 python tools/gpu_conv_probe.py >> gpurun_out/b_probe.jsonl 2>> gpurun_out/b_probe.err
for impl in pair halo; do
  SMB_CONV_IMPL=$impl timeout 300 python tools/gpu_conv_probe.py >> gpurun_out/b_probe.jsonl 2>> gpurun_out/b_probe.err
done
cat gpurun_out/b_probe.jsonl
