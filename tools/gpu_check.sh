#!/bin/bash
# GPU validation driver: every group in its own process (a trap in one tcgen05 kernel must not hide the other
# results), logs under gpurun_out/.  Usage (repo root on the GPU box): bash tools/gpu_check.sh
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, pytest args...
  local name=$1; shift
  echo "=== $name"
  SMB_PARITY_LOG=gpurun_out/parity_stats.jsonl timeout 1200 python -m pytest "$@" -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  grep -E "passed|failed|error|exit" gpurun_out/$name.log | tail -n 4
  grep -E "^(FAILED|ERROR)" gpurun_out/$name.log | head -n 30
}
run texture tests/test_gpu_texture.py
run units_simt tests/test_gpu_vgg_units.py -k "simt or maxpool"
run units_tc1 tests/test_gpu_vgg_units.py -k "tc1"
run units_tc tests/test_gpu_vgg_units.py -k "(tc or pair or halo or ph) and not tc1"
run pipeline_simt tests/test_gpu_pipeline.py -k "simt"
run pipeline_tc tests/test_gpu_pipeline.py -k "not simt"
run fullsize tests/test_gpu_fullsize_properties.py
run cli tests/test_gpu_cli.py
run view tests/test_gpu_view_store.py
run dist tests/test_gpu_dist_adam.py
echo "=== smoke"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
