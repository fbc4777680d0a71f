#!/bin/bash
# round-2 GPU call O: restored conv kernel - whole GPU suite, then same-box A/B: product, SMB_PH_KNOB=1 (halo requested
# at every tap), and a variant whose producer thread executes griddepcontrol.wait / launch_dependents on its own
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02o_pytest_gpu.log
tail -3 gpurun_out/r02o_pytest_gpu.log | cut -c1-200
SMB_PH_KNOB=1 timeout 300 python -m pytest tests/test_gpu_vgg_units.py tests/test_gpu_pipeline.py -q 2>&1 | tail -2
V2=$PWD/stylemesh_b200/lib/libstylemesh_b200_v2.so
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 --no-e2e"
timeout 200 $B > gpurun_out/r02o_bench_product.json 2>/dev/null
SMB_PH_KNOB=1 timeout 200 $B > gpurun_out/r02o_bench_knob1.json 2>/dev/null
SMB_LIB=$V2 timeout 200 $B > gpurun_out/r02o_bench_pdlsplit.json 2>/dev/null
timeout 200 $B > gpurun_out/r02o_bench_product2.json 2>/dev/null
python - <<'PY'
import json
for n in ["product", "knob1", "pdlsplit", "product2"]:
    try:
        d = json.load(open(f"gpurun_out/r02o_bench_{n}.json"))
        k = d["kernel_ms_per_step"]
        print(n, round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "ms  fwd", k["igemm_conv_fwd"], "dgrad", k["igemm_conv_dgrad"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 400 python bench.py > gpurun_out/r02o_bench_c2_n1.json 2>/dev/null; head -c 200 gpurun_out/r02o_bench_c2_n1.json; echo
