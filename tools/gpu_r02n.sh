#!/bin/bash
# round-2 GPU call N: same-box A/B of the conv kernel before (commit 814f6a7) and after the start-up changes
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,pcie.link.gen.current --format=csv > gpurun_out/r02n_gpu.txt 2>&1
python -c "import torch; p=torch.cuda.get_device_properties(0); print(p.name, p.multi_processor_count, p.total_memory, p.L2_cache_size)" >> gpurun_out/r02n_gpu.txt 2>&1
cat gpurun_out/r02n_gpu.txt
O=$PWD/stylemesh_b200/lib/libstylemesh_b200_old.so
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 --no-e2e"
for v in new old new2 old2; do
  case $v in
    new|new2) timeout 200 $B > gpurun_out/r02n_bench_$v.json 2>/dev/null ;;
    old|old2) SMB_LIB=$O timeout 200 $B > gpurun_out/r02n_bench_$v.json 2>/dev/null ;;
  esac
done
python - <<'PY'
import json
for n in ["new", "old", "new2", "old2"]:
    try:
        d = json.load(open(f"gpurun_out/r02n_bench_{n}.json"))
        k = d["kernel_ms_per_step"]
        print(n, round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "ms  fwd", k["igemm_conv_fwd"], "dgrad", k["igemm_conv_dgrad"])
    except Exception as e:
        print(n, "failed", e)
PY
