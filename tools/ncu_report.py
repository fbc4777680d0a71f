"""Turn the ncu outputs of tools/gpu_r02h.sh into the committed evidence (run HERE, no GPU needed):

    python tools/ncu_report.py launches gpurun_out/r02h_launches_ncu.csv profiles/r02h_step_timeline.txt
        -> one steady-state step printed in launch order (duration, grid, kernel) + per-kernel-class shares
    python tools/ncu_report.py full gpurun_out/r02h_full_step.ncu-rep profiles/r02h_ncu_full_step
        -> <out>.json (tools/ncu_summary.py format) and <out>.md (table: duration, DRAM bytes / GB/s / %, L2 %, tensor %)
"""
import csv, io, json, re, subprocess, sys


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("smb::", "")
    return name[:48]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    items = [(int(r[0]), r[4], r[8], float(r[14].replace(",", ""))) for r in rows if r[12] == "gpu__time_duration.sum"]
    unit = next((r[13] for r in rows if r[12] == "gpu__time_duration.sum"), "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3}.get(unit, 1e-3)
    # a steady-state step: from one adam kernel (exclusive) to the next (inclusive), taken late in the run
    adam = [i for i, it in enumerate(items) if "adam_clamp_reg_seg" in it[1]]
    if len(adam) < 3:
        raise SystemExit("not enough steps in the launch list")
    a, b = adam[-3] + 1, adam[-2] + 1
    step = items[a:b]
    with open(out, "w") as fh:
        tot = 0.0
        cls = {}
        for _, name, grid, t in step:
            us = t * scale
            tot += us
            k = short(name)
            cls[k] = cls.get(k, 0.0) + us
            fh.write(f"{us:8.1f} {grid:<16} {k}\n")
        fh.write(f"step total us {tot:.1f} launches {len(step)}   (ncu: cold cache, serialised replays - compare shares)\n\n")
        for k, v in sorted(cls.items(), key=lambda kv: -kv[1]):
            fh.write(f"{v:8.1f} us {100 * v / tot:5.1f} %  {k}\n")
    print(f"{len(step)} launches, {tot:.1f} us -> {out}")


def full(rep, out):
    subprocess.run([sys.executable, __file__.replace("ncu_report.py", "ncu_summary.py"), rep, out + ".json"], check=True)
    rows = json.load(open(out + ".json"))
    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tunit = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}

    def num(txt, table):
        v, u = (txt.split() + [""])[:2]
        return float(v.replace(",", "")) * table.get(u, 1.0)

    with open(out + ".md", "w") as fh:
        fh.write(f"# `ncu --set full --clock-control none` of {len(rows)} consecutive launches of `bench.py` (about one step)\n\n"
                 "Cold-cache, serialised replays: durations are longer than in the running step (compare shares). DRAM GB/s = "
                 "(read + write bytes) / duration; `tensor %` = sm__pipe_tensor_cycles_active (of elapsed).\n\n"
                 "| # | kernel | grid | µs | DRAM MB | DRAM GB/s | DRAM % | L2 % | tensor % |\n|---|---|---|---|---|---|---|---|---|\n")
        for i, r in enumerate(rows):
            us = num(r.get("gpu__time_duration.sum", "0 us"), tunit)
            mb = (num(r.get("dram__bytes_read.sum", "0 byte"), units) + num(r.get("dram__bytes_write.sum", "0 byte"), units)) / 1e6
            pct = lambda k: r.get(k, "0").split()[0]
            fh.write(f"| {i} | `{short(r['Kernel Name'])}` | {r.get('launch__grid_size', '').split()[0] if r.get('launch__grid_size') else ''} "
                     f"| {us:.1f} | {mb:.1f} | {mb / max(us, 1e-9) * 1e3:.0f} | {float(pct('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or 0):.1f} "
                     f"| {float(pct('lts__throughput.avg.pct_of_peak_sustained_elapsed') or 0):.1f} "
                     f"| {float(pct('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed') or 0):.1f} |\n")
    print(f"-> {out}.json, {out}.md")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
