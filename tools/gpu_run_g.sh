#!/bin/bash
# PDL + in-kernel Gram masking validation: unit/pipeline/full-size tests with PDL on, bench A/B (SMB_PDL=0 vs 1)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)" gpurun_out/$name.log | head -n 20; }
run g_units tests/test_gpu_vgg_units.py -k "gram or ph or maxpool"
run g_texture tests/test_gpu_texture.py
run g_pipe tests/test_gpu_pipeline.py -k "not simt"
run g_full tests/test_gpu_fullsize_properties.py
for pdl in 0 1; do
  SMB_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_pdl$pdl.json 2> gpurun_out/g_bench_pdl$pdl.err
  echo "bench pdl=$pdl exit $?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/g_bench_pdl$pdl.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d["e2e"]["value"], d["with_cached_content_targets"]["value"])
PY
done
