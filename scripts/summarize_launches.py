"""Summarise an `ncu --metrics gpu__time_duration.sum` launch list (CSV or ncu's text log): per-kernel totals.

    python scripts/summarize_launches.py launches.csv [top] [first_launch] [last_launch]
first/last select a window of launch indices (e.g. one steady-state step of the bench).
"""
import collections
import csv
import re
import sys

UNIT = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}


def read(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    if lines and lines[0].lstrip().startswith('"ID"'):
        for row in csv.DictReader(lines):
            if row.get("Metric Name") == "gpu__time_duration.sum":
                out.append((row["Kernel Name"], float(row["Metric Value"].replace(",", "")) * UNIT[row["Metric Unit"]]))
        return out
    name = None
    for l in lines:      # text log: a kernel header line, then its metric table
        m = re.match(r"^\s{2}(\S.*?)\s+\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\), Context", l)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"^\s+gpu__time_duration\.sum\s+(\S+)\s+([\d.,]+)", l)
        if m and name is not None:
            out.append((name, float(m.group(2).replace(",", "")) * UNIT[m.group(1)]))
            name = None
    return out


def main(path, top=30, first=0, last=None):
    rows = read(path)[first:last]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, v in rows:
        name = re.sub(r"\(.*", "", name)[:80]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: launches [{first}:{last}] = {sum(v[0] for v in agg.values())}, {tot:.3f} ms total (cold-cache, serialised)")
    print(f"{'ms':>12} {'share':>7} {'n':>5}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:12.3f} {100 * v[1] / tot:6.2f}% {v[0]:5d}  {k}")


if __name__ == "__main__":
    a = sys.argv
    main(a[1], int(a[2]) if len(a) > 2 else 30, int(a[3]) if len(a) > 3 else 0, int(a[4]) if len(a) > 4 else None)
