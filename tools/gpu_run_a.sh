#!/bin/bash
# Round-1 re-entry probe: correctness of the halo / pair conv kernels, per-impl forward timing, bench per impl.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_vgg_units.py -q -m gpu -k "halo or pair" --timeout 600 -p no:cacheprovider > gpurun_out/a_units.log 2>&1
echo "units exit $?"; tail -n 3 gpurun_out/a_units.log
for impl in tc pair halo; do
  SMB_CONV_IMPL=$impl timeout 300 python tools/gpu_conv_probe.py > gpurun_out/a_probe_$impl.json 2> gpurun_out/a_probe_$impl.err
  echo "probe $impl exit $?"; tail -n 1 gpurun_out/a_probe_$impl.json
done
for impl in tc halo; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-impl $impl > gpurun_out/a_bench_$impl.json 2> gpurun_out/a_bench_$impl.err
  echo "bench $impl exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/a_bench_$impl.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d["e2e"]["value"])
except Exception as e:
    print("parse fail", e)
PY
done
