#!/bin/bash
# r01c: tests for the Gram changes, bench, launch list and a full ncu capture of every NON-conv kernel of one step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)" gpurun_out/$name.log | head -n 20; }
run p_units tests/test_gpu_vgg_units.py -k "gram"
run p_pipe tests/test_gpu_pipeline.py -k "not simt"
run p_full tests/test_gpu_fullsize_properties.py
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/p_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d["e2e"]["value"], d["with_cached_content_targets"]["value"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv \
    --log-file gpurun_out/r01c_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01c_launches_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r01c_launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"conv_first_tc|gram_tc|gram_mse|maxpool|igemm_tc2|uv_|adam|content_mse|relu_mask|sumsq" -s 150 -c 45 \
    -o gpurun_out/r01c_small -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01c_small.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/r01c_small.ncu-rep
