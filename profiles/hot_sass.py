"""Top SASS instructions of one kernel of an .ncu-rep by executed count, with their CUDA source line (nvdisasm line info).
usage: hot_sass.py <rep> <kernel regex> <mangled substring> <cubin> [top] [file:lo-hi filter]"""
import csv
import re
import subprocess
import sys

rep, kre, mangled, cubin = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
flt = sys.argv[6] if len(sys.argv) > 6 else None
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
items = []
for r in rows:
    if not r["Address"].startswith("0x"):
        break
    a = int(r["Address"], 16) - base
    f, ln = lines.get(a) or ("?", 0)
    items.append((a, int(r["Instructions Executed"] or 0), int(r["Thread Instructions Executed"] or 0), int(r["Warp Stall Sampling (All Samples)"] or 0), r["Source"], f, ln))
tot = sum(i[1] for i in items)
if flt:
    ff, rng = flt.split(":")
    lo, hi = map(int, rng.split("-"))
    sel = [i for i in items if i[5] == ff and lo <= i[6] <= hi]
    print(f"{flt}: {100 * sum(i[1] for i in sel) / tot:.1f}% of warp instructions")
    for i in sel:
        print(f"  {i[0]:5x} {100*i[1]/tot:5.2f}% lanes {i[2]/max(1,i[1]):4.1f} st {i[3]:4d}  {i[5]}:{i[6]:<4d} {i[4][:90]}")
else:
    for i in sorted(items, key=lambda x: -x[1])[:top]:
        print(f"  {i[0]:5x} {100*i[1]/tot:5.2f}% lanes {i[2]/max(1,i[1]):4.1f} st {i[3]:4d}  {i[5]}:{i[6]:<4d} {i[4][:90]}")
