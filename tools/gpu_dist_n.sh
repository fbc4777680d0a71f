#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dist_adam_check.py > gpurun_out/v_check_n$N.log 2>&1
echo "check exit $?"; grep -E "^\{|watchdog|Error|error" gpurun_out/v_check_n$N.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/v_bench_n${N}.json 2> gpurun_out/v_bench_n${N}.err
echo "bench N=$N exit $?"; grep -E "watchdog|Error|falling back" gpurun_out/v_bench_n${N}.err | head -5
python - <<PY
import json
d = json.loads(open("gpurun_out/v_bench_n${N}.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "host_enqueue_ms_per_step")}, d["e2e"]["value"])
PY
