#!/bin/bash
# compute-sanitizer memcheck over the export / mip-preview kernels
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_texture.py -k "export or mip" -x -q > gpurun_out/r02u_sanitizer_memcheck_export.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r02u_sanitizer_memcheck_export.log | cut -c1-300
