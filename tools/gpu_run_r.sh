#!/bin/bash
# gram A==B aliasing, conv1_1 TMA-store epilogue, segmented Adam / regulariser: tests, bench, launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)|watchdog" gpurun_out/$name.log | head -n 30; }
run r_units tests/test_gpu_vgg_units.py -k "gram or first or fused or maxpool"
run r_texture tests/test_gpu_texture.py
run r_pipe tests/test_gpu_pipeline.py -k "not simt"
run r_full tests/test_gpu_fullsize_properties.py
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["with_cached_content_targets"]["value"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv \
    --log-file gpurun_out/r01d_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01d_launches_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r01d_launches.csv
