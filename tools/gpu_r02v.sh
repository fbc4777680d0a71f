#!/bin/bash
# compute-sanitizer memcheck over the view-preparation kernels (smb_view_*)
mkdir -p gpurun_out
timeout 80 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_view_store.py -x -q > gpurun_out/r02v_sanitizer_memcheck_view_store.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r02v_sanitizer_memcheck_view_store.log | cut -c1-300
