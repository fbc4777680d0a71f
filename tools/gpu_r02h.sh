#!/bin/bash
# round-2 GPU call H (1 GPU): final state - whole GPU suite, smoke, bench (both arms), ncu launch list + full capture
set -u
mkdir -p gpurun_out
export SMB_PARITY_LOG=gpurun_out/r02h_parity_stats.jsonl
rm -f $SMB_PARITY_LOG
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02h_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest_gpu.log
tail -12 gpurun_out/r02h_pytest_gpu.log
unset SMB_PARITY_LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02h_smoke.log
timeout 600 python bench.py > gpurun_out/r02h_bench_c2_n1.json 2> gpurun_out/r02h_bench_c2_n1.err
echo "bench rc=$?"; head -c 300 gpurun_out/r02h_bench_c2_n1.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02h_bench_reference.json 2>/dev/null
echo "reference rc=$?"; head -c 300 gpurun_out/r02h_bench_reference.json; echo
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --sustained-s 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_launches_ncu.csv $B > /dev/null 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/r02h_launches_ncu.csv
timeout 900 ncu --set full --clock-control none --import-source on -s 700 -c 64 -f -o gpurun_out/r02h_full_step $B > gpurun_out/r02h_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r02h_full_step.ncu-rep
