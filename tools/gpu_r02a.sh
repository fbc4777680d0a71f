#!/bin/bash
# round-2 GPU call A: full-size parity, C2/C3/C4 bench lines at N=1, sanitizer logs (single GPU)
set -u
mkdir -p gpurun_out
export SMB_PARITY_LOG=gpurun_out/r02a_parity_stats.jsonl
rm -f $SMB_PARITY_LOG
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize_parity.py -x -q > gpurun_out/r02a_pytest_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r02a_pytest_fullsize.log
tail -5 gpurun_out/r02a_pytest_fullsize.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_c2_n1.json 2> gpurun_out/r02a_bench_c2_n1.err
echo "bench c2 rc=$?"; head -c 600 gpurun_out/r02a_bench_c2_n1.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --preset with_angle_and_depth --view 256x341 > gpurun_out/r02a_bench_c3_n1.json 2> gpurun_out/r02a_bench_c3_n1.err
echo "bench c3 rc=$?"; head -c 400 gpurun_out/r02a_bench_c3_n1.json; echo
timeout 900 python bench.py --steps 20 --warmup 3 --preset with_angle_and_depth --view 256x320 --texture 4096 > gpurun_out/r02a_bench_c4_n1.json 2> gpurun_out/r02a_bench_c4_n1.err
echo "bench c4 rc=$?"; head -c 400 gpurun_out/r02a_bench_c4_n1.json; echo
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r02a_sanitizer_memcheck_smoke.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_sanitizer_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r02a_sanitizer_racecheck_smoke.log
