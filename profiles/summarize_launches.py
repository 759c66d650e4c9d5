"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into a
per-kernel table (count, total ms, share). Usage: python profiles/summarize_launches.py launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    rows = [r for r in rows if r[0] >= skip]
    tot = defaultdict(lambda: [0, 0.0])
    for _, name, ns in rows:
        short = re.sub(r"\(.*", "", name).replace("<unnamed>::", "").replace("void ", "")
        short = re.sub(r"cub::(\w+)<.*", r"cub::\1", short)
        tot[short][0] += 1
        tot[short][1] += ns
    total = sum(v[1] for v in tot.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% |")
    print(f"| **total** | {len(rows)} | {total / 1e6:.3f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
