#!/bin/bash
# final validation of the tree as committed: smoke, the GPU suite, the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r02p_bench_c2_n1.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
cat gpurun_out/r02p_bench_c2_n1.json | cut -c1-400
