"""Print one step of an ncu launch list (gpu__time_duration.sum CSV): python tools/launch_timeline.py csv [step_index]"""
import csv, sys


def main(path, step=1):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    L = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[hdr + 2:] if len(r) > vi]
    idx = [i for i, (n, t, g) in enumerate(L) if "uv_sample" in n]
    a, b = idx[step], idx[step + 1]
    tot = 0.0
    for n, t, g in L[a:b]:
        tot += t
        print("%8.1f %-14s %s" % (t / 1e3, g, n.replace("void ", "").replace("smb::", "")[:60]))
    print("step total us %.1f launches %d" % (tot / 1e3, b - a))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
