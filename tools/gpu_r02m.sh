#!/bin/bash
# round-2 GPU call M (1 GPU): final tree - whole GPU suite, smoke, default bench + reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02m_pytest_gpu.log
tail -4 gpurun_out/r02m_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02m_smoke.log
timeout 600 python bench.py > gpurun_out/r02m_bench_c2_n1.json 2> gpurun_out/r02m_bench_c2_n1.err
echo "bench rc=$?"; head -c 260 gpurun_out/r02m_bench_c2_n1.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02m_bench_reference.json 2>/dev/null; head -c 200 gpurun_out/r02m_bench_reference.json; echo
