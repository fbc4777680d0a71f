#!/bin/bash
# round-2 GPU call I: where do the N=128 layers wait?  halo-vs-B split of the MMA thread's operand waits, and the
# effect of requesting the halo as three 6-row boxes (SMB_PH_HALO_SPLIT=1)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02i_trace_halo_split.jsonl
rm -f $OUT
for hs in 0 1; do
  for layer in 3 5 9; do
    SMB_PH_HALO_SPLIT=$hs PROBE_LAYER=$layer SMB_CONV_IMPL=ph timeout 120 python tools/gpu_trace_probe.py 2>/dev/null | tail -1 | sed "s/^/{\"halo_split\": $hs, \"probe\": /; s/$/}/" >> $OUT
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02i_trace_halo_split.jsonl'):
    d=json.loads(l); s=d['probe']['summary']
    print('split', d['halo_split'], 'layer', d['probe']['layer'], {k: (int(s[k]['med']), int(s[k]['max'])) for k in ['cycles','mma_first','mma_end','epi_end','w_full','w_full_halo','w_tmem_full','w_tmem_empty']})
PY
timeout 300 python -m pytest tests/test_gpu_texture.py -q -k "export or mip" 2>&1 | tail -6 | cut -c1-300
SMB_PH_HALO_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_vgg_units.py -q -k "ph or fused" 2>&1 | tail -3
SMB_PH_HALO_SPLIT=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r02i_bench_c2_halosplit1.json 2>gpurun_out/r02i_bench_c2_halosplit1.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 > gpurun_out/r02i_bench_c2_halosplit0.json 2>/dev/null
python - <<'PY'
import json
for n in ["halosplit0", "halosplit1"]:
    try:
        d = json.load(open(f"gpurun_out/r02i_bench_c2_{n}.json"))
        print(n, round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "ms", d["kernel_ms_per_step"], d["parity_at_bench_config"]["loss_rel_err"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02i_bench_c2_halosplit1.err
timeout 600 python bench.py > gpurun_out/r02i_bench_c2_n1.json 2> gpurun_out/r02i_bench_c2_n1.err
echo "bench rc=$?"; head -c 200 gpurun_out/r02i_bench_c2_n1.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02i_bench_reference.json 2>/dev/null
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --sustained-s 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_launches_ncu.csv $B > /dev/null 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/r02i_launches_ncu.csv; du -sh gpurun_out | tail -1
