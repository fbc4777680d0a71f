#!/bin/bash
# last check of the round: the default bench line with the final bench.py
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r02s_bench_c2_n1.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02s_bench_c2_n1.json"))
print(round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "sustained", round(d["sustained"]["value"], 1),
      "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"],
      "parity", d["parity_at_bench_config"].get("loss_rel_err"), "cpu", round(d["cpu_baseline"]["value"], 2))
PY
