#!/bin/bash
# Profiling pass for profiles/ (B200_PROFILING.md recipe).  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r01b}
# 1. every launch of a short bench run with its device time (cold-cache, serialised: compare shares)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/${TAG}_launches.csv
# 2. full capture of the dominant kernel (pair + halo conv): 10 launches of a steady-state forward
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:igemm_ph -s 24 -c 12 \
    -o gpurun_out/${TAG}_ph -f python tools/gpu_conv_probe.py > gpurun_out/${TAG}_ph.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/${TAG}_ph.ncu-rep
