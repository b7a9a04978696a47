"""Attribute an ncu SASS source page (ncu -i rep --page source --csv --kernel-name regex:K) to CUDA source lines using
nvdisasm -g line info of the cubin.  usage: hot_lines.py <rep> <kernel regex> <mangled substring> <cubin>"""
import collections
import csv
import re
import subprocess
import sys

rep, kre, mangled, cubin = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
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
        break  # next kernel instance
    key = lines.get(int(r["Address"], 16) - base, ("?", 0))
    v = [int(r["Instructions Executed"] or 0), int(r["Thread Instructions Executed"] or 0), int(r["Warp Stall Sampling (All Samples)"] or 0)]
    for k in range(3):
        agg[key][k] += v[k]; tot[k] += v[k]
print(f"{kre}: warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  stall samples {tot[2]:,}")
src_cache = {}
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    f, ln = key
    text = ""
    for cand in ("phyngsc_b200/csrc/" + f,):  # noqa
        try:
            src_cache.setdefault(cand, open(cand).read().splitlines())
            text = src_cache[cand][ln - 1].strip()[:90]
        except Exception:  # noqa: BLE001
            pass
    print(f"  {f}:{ln:<5d} winst {100*v[0]/max(1,tot[0]):5.1f}%  samples {100*v[2]/max(1,tot[2]):5.1f}%   {text}")
