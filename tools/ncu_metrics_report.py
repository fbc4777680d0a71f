"""Turn the long-format CSV of an `ncu --metrics ... --csv --log-file` pass (tools/gpu_r02r.sh) into the committed
evidence (run HERE, no GPU needed):

    python tools/ncu_metrics_report.py gpurun_out/r02r_ncu_metrics.csv profiles/r02r_ncu_metrics_step

-> <out>.json (tools/ncu_summary.py's format: one dict per launch, "value unit" strings; bench.py reads the igemm_ph
DRAM bytes for `roofline.traffic` from it) and <out>.md (one steady-state step as a table + per-class sums).
"""
import csv, json, re, sys


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("smb::", "")
    return name[:44]


def main(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"Kernel Name": r[4][:160], "ID": r[0], "launch__grid_size": r[8],
                                            "launch__block_size": r[7]})
        d[r[12]] = f"{r[14]} {r[13]}".strip()
    items = [launches[k] for k in sorted(launches)]
    # one steady-state step: from one optimiser launch (exclusive) to the next (inclusive)
    adam = [i for i, it in enumerate(items) if "adam_clamp_reg_seg" in it["Kernel Name"]]
    if len(adam) < 2:
        raise SystemExit("fewer than two optimiser launches in the capture")
    step = items[adam[0] + 1:adam[1] + 1]
    json.dump(step, open(out + ".json", "w"), indent=1)

    def val(d, k):
        return float(d.get(k, "0").split()[0].replace(",", ""))

    tunit = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
    with open(out + ".md", "w") as fh:
        fh.write(f"# ncu metrics pass (`--clock-control none`) over one steady-state step of `bench.py` ({len(step)} launches)\n\n"
                 "Serialised, cold-cache replays: durations are longer than in the running step (compare shares). DRAM MB = "
                 "dram__bytes_read.sum + dram__bytes_write.sum; tensor % = sm__pipe_tensor_cycles_active (of elapsed); L2 hit = "
                 "lts__t_sector_hit_rate.\n\n| # | kernel | grid | µs | DRAM MB | DRAM GB/s | DRAM % | L2 % | L2 hit % | tensor % | regs |\n"
                 "|---|---|---|---|---|---|---|---|---|---|---|\n")
        cls, tot_us, tot_mb = {}, 0.0, 0.0
        for i, d in enumerate(step):
            tu = d.get("gpu__time_duration.sum", "0 ns").split()
            us = float(tu[0].replace(",", "")) * tunit.get(tu[1] if len(tu) > 1 else "ns", 1e-3)
            mb = (val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")) / 1e6
            k = short(d["Kernel Name"])
            c = cls.setdefault(k, [0, 0.0, 0.0])
            c[0] += 1; c[1] += us; c[2] += mb
            tot_us += us; tot_mb += mb
            fh.write(f"| {i} | `{k}` | {d['launch__grid_size']} | {us:.1f} | {mb:.1f} | {mb / us * 1e3 if us else 0:.0f} | "
                     f"{val(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{val(d, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {val(d, 'lts__t_sector_hit_rate.pct'):.1f} | "
                     f"{val(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{val(d, 'launch__registers_per_thread'):.0f} |\n")
        fh.write(f"\nStep total: {tot_us:.1f} µs over {len(step)} launches, {tot_mb:.0f} MB of DRAM traffic.\n\n"
                 "| kernel | launches | µs | share | DRAM MB |\n|---|---|---|---|---|\n")
        for k, (n, us, mb) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot_us:.1f} % | {mb:.1f} |\n")
    print(f"{len(step)} launches, {tot_us:.1f} us, {tot_mb:.0f} MB -> {out}.json / .md")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
