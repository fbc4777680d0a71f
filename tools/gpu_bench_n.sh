#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "bench N=$N exit $?"; tail -n 2 gpurun_out/n${N}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/n${N}_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "kernel_ms_per_step")}, d["e2e"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/n${N}_ref.json 2> gpurun_out/n${N}_ref.err
echo "ref exit $?"; cat gpurun_out/n${N}_ref.json | cut -c1-300
