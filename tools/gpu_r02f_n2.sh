#!/bin/bash
# round-2 GPU call F (2 GPUs): bench sanity at N=2 after the collective-setup fix, dist tests with the new bars
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --sustained-s 2 > gpurun_out/r02f_bench_c2_n2.json 2> gpurun_out/r02f_bench_c2_n2.err
echo "bench n2 rc=$?"; head -c 250 gpurun_out/r02f_bench_c2_n2.json; echo; tail -3 gpurun_out/r02f_bench_c2_n2.err
timeout 300 python -m pytest tests/test_gpu_dist_adam.py -q > gpurun_out/r02f_pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -3 gpurun_out/r02f_pytest_dist.log
