#!/bin/bash
# round-2 GPU call K (2 GPUs): the whole GPU suite on the final code, 2-rank tests included; smoke
set -u
mkdir -p gpurun_out
export SMB_PARITY_LOG=gpurun_out/r02k_parity_stats.jsonl
rm -f $SMB_PARITY_LOG
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02k_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02k_pytest_gpu.log
tail -14 gpurun_out/r02k_pytest_gpu.log | cut -c1-300
unset SMB_PARITY_LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02k_smoke.log
