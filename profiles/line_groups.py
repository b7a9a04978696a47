"""Group an ncu source page by source-line ranges.  usage: line_groups.py <rep> <kernel regex> <mangled substring> <cubin>
Prints warp-instruction share and stall-sample share per function-sized group of phy_core.cuh / phy_kernels.cuh."""
import collections
import csv
import re
import subprocess
import sys

rep, kre, mangled, cubin = sys.argv[1:5]
src = {}
for f in ("phy_core.cuh", "phy_kernels.cuh", "phy_encode.cuh", "phy_fast.cuh", "phy_seqstat.cuh"):
    src[f] = open("phyngsc_b200/csrc/" + f).read().splitlines()


def group_of(f, ln):
    """name of the enclosing top-level definition (struct / function) of line ln"""
    if f not in src:
        return f
    L = src[f]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        t = L[i]
        if t and not t[0].isspace() and not t.startswith(("}", "/", "*", "#")) and ("(" in t or t.startswith("struct")):
            m = re.search(r"(\w+)\s*\(", t) if not t.startswith("struct") else re.search(r"struct\s+(\w+)", t)
            return f.split(".")[0][4:] + ":" + (m.group(1) if m else t[:30])
    return f


dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = {}, None, False
for l in dis:
    if l.startswith("\t.section"):
        on = mangled in l and ".text." in l
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        lines[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
rows = list(csv.DictReader(out[start:]))
base = int(rows[0]["Address"], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows:
    if not r["Address"].startswith("0x"):
        break
    f, ln = lines.get(int(r["Address"], 16) - base, ("?", 0)) or ("?", 0)
    key = group_of(f, ln)
    v = [int(r["Instructions Executed"] or 0), int(r["Thread Instructions Executed"] or 0), int(r["Warp Stall Sampling (All Samples)"] or 0)]
    for k in range(3):
        agg[key][k] += v[k]; tot[k] += v[k]
print(f"{kre}: warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  lanes {tot[1]/max(1,tot[0]):.1f}  stall samples {tot[2]:,}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] * 200 < tot[0] and v[2] * 200 < tot[2]:
        continue
    print(f"  {key:34s} winst {100*v[0]/max(1,tot[0]):5.1f}%  lanes {v[1]/max(1,v[0]):4.1f}  samples {100*v[2]/max(1,tot[2]):5.1f}%")
