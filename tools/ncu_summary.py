"""Summarise an .ncu-rep (read here, no GPU needed) into JSON for profiles/:  python tools/ncu_summary.py rep out.json"""
import csv, io, json, subprocess, sys

KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        item = {"Kernel Name": d.get("Kernel Name", "")[:160], "ID": d.get("ID")}
        for k in KEEP:
            if k in d:
                item[k] = f"{d[k]} {u.get(k, '')}".strip()
        res.append(item)
    json.dump(res, open(out, "w"), indent=1)
    print(f"{len(res)} launches -> {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
