#!/bin/bash
# round-2 GPU call J: start-up of the conv launches - weight tiles requested before griddepcontrol.wait (default) vs
# after it (SMB_PH_KNOB=32); time stamps of the dependency wait, the first halo and the first MMA per CTA
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02j_trace_startup.jsonl
rm -f $OUT
for knob in 32 0; do
  for layer in 1 5 9; do
    SMB_PH_KNOB=$knob PROBE_LAYER=$layer SMB_CONV_IMPL=ph timeout 120 python tools/gpu_trace_probe.py 2>/dev/null | tail -1 | sed "s/^/{\"knob\": $knob, \"probe\": /; s/$/}/" >> $OUT
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02j_trace_startup.jsonl'):
    d=json.loads(l); s=d['probe']['summary']
    print('knob', d['knob'], 'layer', d['probe']['layer'], {k: (int(s[k]['med']), int(s[k]['max'])) for k in ['cycles','prologue','dep_wait_done','first_halo','mma_first','mma_end','epi_end','w_full','w_full_halo']})
PY
timeout 900 python -m pytest tests/test_gpu_vgg_units.py tests/test_gpu_pipeline.py -q 2>&1 | tail -3
SMB_PH_KNOB=32 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 --no-e2e > gpurun_out/r02j_bench_c2_noprefetch.json 2>/dev/null
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-s 0 --no-e2e > gpurun_out/r02j_bench_c2_prefetch.json 2>gpurun_out/r02j_bench_c2_prefetch.err
python - <<'PY'
import json
for n in ["noprefetch", "prefetch"]:
    try:
        d = json.load(open(f"gpurun_out/r02j_bench_c2_{n}.json"))
        print(n, round(d["value"], 1), "views/s", round(d["ms_per_step"], 4), "ms", d["kernel_ms_per_step"], d["parity_at_bench_config"]["loss_rel_err"]["total"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02j_bench_c2_prefetch.err
