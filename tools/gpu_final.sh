#!/bin/bash
# Round-end verification + evidence: full GPU suite, smoke, bench (both arms), launch list, ncu --set full of one step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash tools/gpu_check.sh
timeout 900 python bench.py > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench_n1.err
echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r01f_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["e2e"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["roofline_hbm"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01f_bench_reference.json 2> gpurun_out/r01f_bench_reference.err
echo "reference exit $?"; cut -c1-200 gpurun_out/r01f_bench_reference.json
timeout 300 python tools/view_store_bench.py > gpurun_out/r01f_view_store.json 2> gpurun_out/r01f_view_store.err; cat gpurun_out/r01f_view_store.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv \
    --log-file gpurun_out/r01f_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01f_launches_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r01f_launches.csv
timeout 1200 ncu --set full --clock-control none -s 300 -c 72 -o gpurun_out/r01f_step -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01f_step.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/r01f_step.ncu-rep
python tools/ncu_summary.py gpurun_out/r01f_step.ncu-rep gpurun_out/r01f_ncu_full_step.json
sz=$(stat -c %s gpurun_out/r01f_step.ncu-rep); if [ "$sz" -gt 40000000 ]; then rm gpurun_out/r01f_step.ncu-rep; echo "rep too large to pull ($sz), summary kept"; fi
