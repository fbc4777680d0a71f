#!/bin/bash
# round-2 GPU call L: two cheap A/Bs on one box - early halo request (SMB_PH_KNOB=1) and the WIDE64 build of igemm_ph<64>
set -u
mkdir -p gpurun_out
W=$PWD/stylemesh_b200/lib/libstylemesh_b200_wide64.so
SMB_LIB=$W timeout 300 python -m pytest tests/test_gpu_vgg_units.py tests/test_gpu_fullsize_parity.py -q -k "ph or fused or benchmark_shapes or C2" 2>&1 | tail -3
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 --no-e2e"
for v in base knob1 wide64 base2; do
  case $v in
    base|base2) timeout 200 $B > gpurun_out/r02l_bench_$v.json 2>/dev/null ;;
    knob1) SMB_PH_KNOB=1 timeout 200 $B > gpurun_out/r02l_bench_$v.json 2>/dev/null ;;
    wide64) SMB_LIB=$W timeout 200 $B > gpurun_out/r02l_bench_$v.json 2>/dev/null ;;
  esac
done
python - <<'PY'
import json
for n in ["base", "knob1", "wide64", "base2"]:
    try:
        d = json.load(open(f"gpurun_out/r02l_bench_{n}.json"))
        k = d["kernel_ms_per_step"]
        print(n, round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "ms  fwd", k["igemm_conv_fwd"], "dgrad", k["igemm_conv_dgrad"], "parity", d["parity_at_bench_config"]["loss_rel_err"]["total"])
    except Exception as e:
        print(n, "failed", e)
PY
