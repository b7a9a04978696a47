"""profiles/ncu_traffic.json from a full capture's metric dump (profiles/<tag>_ncu_full_metrics_<shape>.txt, written by
ncu_summary.py): DRAM bytes read + written per launch of every record kernel on the captured 256 MB shard, beside the
kernel's algorithmic bytes on the same shard (the formulas bench.py uses for roofline.achieved).
usage: python profiles/make_traffic.py <tag> [shape] [out/in ratio]"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyngsc_b200 import synth  # noqa: E402
from bench import record_stats  # noqa: E402

tag = sys.argv[1]
shape = sys.argv[2] if len(sys.argv) > 2 else "100bp"
ratio = float(sys.argv[3]) if len(sys.argv) > 3 else 0.358
data = synth.fastq(shape, 2, target_bytes=256 * 1_000_000)  # the shard tests/gpu_prof_target.py compresses
nrec, title_b, seq_b = record_stats(data)
bytes_in, bytes_out = data.size, data.size * ratio
alg = {"nl_count": bytes_in + bytes_in / 8, "nl_emit": bytes_in / 8 + 12 * nrec, "stat1": title_b + 4 * nrec, "seqstat": 2 * seq_b + 2 * nrec,
       "stat2": title_b + 8 * nrec, "enc_title": title_b + 0.12 * bytes_out, "enc_qd": 2 * seq_b + 0.88 * bytes_out, "place": 2 * bytes_out}
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {"note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture (profiles/{tag}_ncu_full_metrics_{shape}.txt; "
               f"PHY_GROUPS=1 so that a launch covers the whole shard), 256 MB {shape} shard; 'algorithmic' = the stage's must-read + must-write bytes on "
               "that shard (DESIGN.md section 3, the same formulas as bench.py)", "shape": shape}
cur = None
for line in open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{tag}_ncu_full_metrics_{shape}.txt")):
    if not line.startswith(" "):
        cur = re.sub(r"^void ", "", line.strip()).split("<")[0]
        continue
    m = re.match(r"\s+(dram_rd|dram_wr)\s+([0-9.,]+)\s+(\w+)", line)
    if m and cur and cur.startswith("k_") and cur[2:] in alg:
        e = res.setdefault(cur, {"shard_bytes": int(bytes_in), "traffic": 0, "algorithmic": int(alg[cur[2:]])})
        e["traffic"] += int(float(m.group(2).replace(",", "")) * unit[m.group(3)])
for k, e in res.items():
    if isinstance(e, dict):
        e["traffic_over_algorithmic"] = round(e["traffic"] / max(1, e["algorithmic"]), 3)
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
