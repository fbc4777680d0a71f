#!/bin/bash
# timing of the "next" rows (f3 export / preview, f4 rasteriser, K5 view plan) with their CPU chains beside them
mkdir -p gpurun_out
timeout 400 python tools/next_rows_bench.py --rows ${ROWS:-f3,f4,k5} > gpurun_out/r02q_next_rows${SUFFIX:-}.jsonl 2> gpurun_out/r02q_next_rows.err; echo "rc=$?"
cat gpurun_out/r02q_next_rows${SUFFIX:-}.jsonl | cut -c1-1500; tail -5 gpurun_out/r02q_next_rows.err
