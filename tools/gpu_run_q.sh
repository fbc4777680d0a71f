#!/bin/bash
# fused Gram backward in the pair + halo dgrad conv: unit tests, pipeline parity, bench A/B (SMB_PH_FUSE=0 vs 1)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)|watchdog" gpurun_out/$name.log | head -n 30; }
run q_fused tests/test_gpu_vgg_units.py -k "fused"
run q_units tests/test_gpu_vgg_units.py -k "ph and not fused"
run q_pipe tests/test_gpu_pipeline.py -k "not simt"
run q_full tests/test_gpu_fullsize_properties.py
for fuse in 0 1; do
  SMB_PH_FUSE=$fuse timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_fuse$fuse.json 2> gpurun_out/q_bench_fuse$fuse.err
  echo "bench fuse=$fuse exit $?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/q_bench_fuse$fuse.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d["e2e"]["value"], d["with_cached_content_targets"]["value"])
PY
done
