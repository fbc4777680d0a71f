#!/bin/bash
# round-2 GPU call C: conv1_2 bottleneck elimination (trace probe x knobs), host profile, rest of the GPU suite
set -u
mkdir -p gpurun_out
for knob in 0 2 4 8 6 10 12 14; do
  SMB_PH_KNOB=$knob PROBE_LAYER=1 SMB_CONV_IMPL=ph timeout 120 python tools/gpu_trace_probe.py 2>/dev/null | tail -1 | sed "s/^/{\"knob\": $knob, \"probe\": /; s/$/}/" >> gpurun_out/r02c_trace_conv1_2_knobs.jsonl
done
SMB_PH_DIRECT_STORES=1 PROBE_LAYER=1 SMB_CONV_IMPL=ph timeout 120 python tools/gpu_trace_probe.py 2>/dev/null | tail -1 | sed "s/^/{\"knob\": \"direct_stores\", \"probe\": /; s/$/}/" >> gpurun_out/r02c_trace_conv1_2_knobs.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02c_trace_conv1_2_knobs.jsonl'):
    d=json.loads(l); s=d['probe']['summary']
    print(d['knob'], {k: s[k]['med'] for k in ['cycles','prologue','mma_first','tma_end','mma_end','epi_first','epi_end','w_tmem_full','w_full','w_tmem_empty','w_empty']})
PY
timeout 300 python tools/host_profile.py > gpurun_out/r02c_host_profile_c2.txt 2>gpurun_out/r02c_host_profile_c2.err; head -40 gpurun_out/r02c_host_profile_c2.txt
timeout 300 python tools/host_profile.py --preset with_angle_and_depth --view 256x341 > gpurun_out/r02c_host_profile_c3.txt 2>gpurun_out/r02c_host_profile_c3.err; head -45 gpurun_out/r02c_host_profile_c3.txt
export SMB_PARITY_LOG=gpurun_out/r02c_parity_stats.jsonl
rm -f $SMB_PARITY_LOG
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -18 gpurun_out/r02c_pytest_gpu.log
