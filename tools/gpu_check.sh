#!/bin/bash
# First-contact GPU validation: every test file in its own process (a trap in one tcgen05 kernel must not hide the
# other results), logs under gpurun_out/.  Usage (from the repo root on the GPU box): bash tools/gpu_check.sh
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_texture test_gpu_vgg_units test_gpu_pipeline; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 ${PYTEST_EXTRA} > gpurun_out/$f.log 2>&1
  echo "exit $?" >> gpurun_out/$f.log
  tail -n 25 gpurun_out/$f.log
done
