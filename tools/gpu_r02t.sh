#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels that had no sanitizer record yet: rasteriser and mask pyramid
mkdir -p gpurun_out
timeout 160 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_raster.py tests/test_gpu_mask_plan.py -k "matches_the_oracle or level_masks or plain_mask" -x -q > gpurun_out/r02t_sanitizer_memcheck_raster_maskplan.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r02t_sanitizer_memcheck_raster_maskplan.log | cut -c1-300
