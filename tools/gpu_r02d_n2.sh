#!/bin/bash
# round-2 GPU call D (2 GPUs): N-rank end-to-end parity with the real kernels at full size, sanitizer on a 2-rank step
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
OUT=gpurun_out/r02d_dist_pipeline_check.jsonl
timeout 300 $TR --nproc-per-node 2 --master-port 29551 tools/dist_adam_check.py 2>gpurun_out/r02d_n2_adam.err | grep '^{' >> gpurun_out/r02d_dist_adam_check.jsonl
echo "adam n2 rc=${PIPESTATUS[0]}"
timeout 600 $TR --nproc-per-node 2 --master-port 29552 tools/dist_pipeline_check.py --preset only2D --view 480x640 --texture 2048 --steps 3 --out $OUT > gpurun_out/r02d_n2_c2.log 2>&1
echo "pipeline n2 C2 rc=$?"; tail -1 gpurun_out/r02d_n2_c2.log | cut -c1-300
timeout 900 $TR --nproc-per-node 2 --master-port 29553 tools/dist_pipeline_check.py --preset with_angle_and_depth --view 256x341 --texture 2048 --steps 2 --out $OUT > gpurun_out/r02d_n2_c3.log 2>&1
echo "pipeline n2 C3 rc=$?"; tail -1 gpurun_out/r02d_n2_c3.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_dist_adam.py -q > gpurun_out/r02d_pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -3 gpurun_out/r02d_pytest_dist.log
timeout 900 compute-sanitizer --tool memcheck --target-processes all --print-limit 10 $TR --nproc-per-node 2 --master-port 29554 tools/dist_pipeline_check.py --preset only2D --view 96x128 --texture 256 --steps 2 > gpurun_out/r02d_sanitizer_memcheck_n2_step.log 2>&1
echo "memcheck n2 rc=$?"; grep -E "ERROR SUMMARY|dist_pipeline_summary" gpurun_out/r02d_sanitizer_memcheck_n2_step.log | tail -6
timeout 600 $TR --nproc-per-node 2 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02d_bench_c2_n2.json 2> gpurun_out/r02d_bench_c2_n2.err
echo "bench n2 rc=$?"; head -c 300 gpurun_out/r02d_bench_c2_n2.json; echo
