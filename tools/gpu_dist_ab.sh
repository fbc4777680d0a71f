#!/bin/bash
# 2-GPU: fused reduce-scatter + Adam + all-gather kernel: unit check, then bench A/B against the NCCL path
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dist_adam_check.py > gpurun_out/u_check.log 2>&1
echo "check exit $?"; grep -E "^\{|watchdog|Error|error" gpurun_out/u_check.log | head -20
for fused in 0 1; do
  SMB_DIST_ADAM=$fused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$fused bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/u_bench_n${N}_fused$fused.json 2> gpurun_out/u_bench_n${N}_fused$fused.err
  echo "bench N=$N fused=$fused exit $?"; grep -E "watchdog|Error|falling back" gpurun_out/u_bench_n${N}_fused$fused.err | head -5
  python - <<PY
import json
d = json.loads(open("gpurun_out/u_bench_n${N}_fused$fused.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "host_enqueue_ms_per_step")}, d["e2e"]["value"])
PY
done
