"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, mean, total, share."""
import collections
import csv
import sys


def main(path, title, first=None):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    seen = 0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        seen += 1
        if first is not None and seen > first:
            break
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
        agg.setdefault(r["Kernel Name"].split("(")[0], []).append(float(r["Metric Value"].replace(",", "")) * scale)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {title}")
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES with bench.py's stage_ms, not absolutes")
    print(f"{'kernel':14s} {'launches':>8s} {'mean_ms':>9s} {'total_ms':>9s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:14s} {len(v):8d} {sum(v) / len(v):9.4f} {sum(v):9.3f} {100 * sum(v) / tot:6.1f}%")


if __name__ == "__main__":
    # optional third argument: only the first N launches (bench.py's kernel-only region: (warmup + steps) x launches per step)
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else None)
