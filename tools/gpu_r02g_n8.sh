#!/bin/bash
# round-2 GPU call G (8 GPUs): N-rank parity at N = 8 (and 4), bench lines of BASELINE configs C2 / C3 / C4 at N = 8
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
OUT=gpurun_out/r02g_dist_pipeline_check.jsonl
timeout 200 $TR --nproc-per-node 8 --master-port 29571 tools/dist_adam_check.py 2>gpurun_out/r02g_n8_adam.err | grep '^{' >> gpurun_out/r02g_dist_adam_check.jsonl
echo "adam n8 rc=${PIPESTATUS[0]}"; tail -1 gpurun_out/r02g_dist_adam_check.jsonl
timeout 400 $TR --nproc-per-node 8 --master-port 29572 tools/dist_pipeline_check.py --preset only2D --view 480x640 --texture 2048 --steps 2 --out $OUT > gpurun_out/r02g_n8_c2.log 2>&1
echo "pipeline n8 C2 rc=$?"; tail -1 gpurun_out/r02g_n8_c2.log | cut -c1-300
timeout 300 $TR --nproc-per-node 4 --master-port 29573 tools/dist_pipeline_check.py --preset with_angle_and_depth --view 128x171 --texture 1024 --steps 2 --out $OUT > gpurun_out/r02g_n4_c3small.log 2>&1
echo "pipeline n4 C3-small rc=$?"; tail -1 gpurun_out/r02g_n4_c3small.log | cut -c1-300
timeout 240 $TR --nproc-per-node 8 --master-port 29574 bench.py --gpus 8 --steps 20 --warmup 3 --sustained-s 2 > gpurun_out/r02g_bench_c2_n8.json 2> gpurun_out/r02g_bench_c2_n8.err
echo "bench c2 n8 rc=$?"; head -c 200 gpurun_out/r02g_bench_c2_n8.json; echo
timeout 300 $TR --nproc-per-node 8 --master-port 29575 bench.py --gpus 8 --steps 20 --warmup 3 --sustained-s 2 --preset with_angle_and_depth --view 256x341 > gpurun_out/r02g_bench_c3_n8.json 2> gpurun_out/r02g_bench_c3_n8.err
echo "bench c3 n8 rc=$?"; head -c 200 gpurun_out/r02g_bench_c3_n8.json; echo
timeout 300 $TR --nproc-per-node 8 --master-port 29576 bench.py --gpus 8 --steps 20 --warmup 3 --sustained-s 2 --preset with_angle_and_depth --view 256x320 --texture 4096 > gpurun_out/r02g_bench_c4_n8.json 2> gpurun_out/r02g_bench_c4_n8.err
echo "bench c4 n8 rc=$?"; head -c 200 gpurun_out/r02g_bench_c4_n8.json; echo
