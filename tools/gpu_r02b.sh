#!/bin/bash
# round-2 GPU call B: whole GPU suite (new full-size parity + checkpoint tests), C3/C4 bench lines at N=1
set -u
mkdir -p gpurun_out
export SMB_PARITY_LOG=gpurun_out/r02b_parity_stats.jsonl
rm -f $SMB_PARITY_LOG
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -25 gpurun_out/r02b_pytest_gpu.log
unset SMB_PARITY_LOG
timeout 600 python bench.py --steps 20 --warmup 3 --preset with_angle_and_depth --view 256x341 > gpurun_out/r02b_bench_c3_n1.json 2> gpurun_out/r02b_bench_c3_n1.err
echo "bench c3 rc=$?"; head -c 300 gpurun_out/r02b_bench_c3_n1.json; echo
timeout 900 python bench.py --steps 20 --warmup 3 --preset with_angle_and_depth --view 256x320 --texture 4096 > gpurun_out/r02b_bench_c4_n1.json 2> gpurun_out/r02b_bench_c4_n1.err
echo "bench c4 rc=$?"; head -c 300 gpurun_out/r02b_bench_c4_n1.json; echo
