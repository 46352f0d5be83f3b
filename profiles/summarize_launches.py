#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: avg per kernel + share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        agg[name].append(v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3))
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k[:64]:64s} n={len(v):4d} avg={sum(v) / len(v):8.2f} us  share={sum(v) / tot:.3f}")
    print(f"sum of averages per layer: {sum(sum(v) / len(v) for v in agg.values()):.2f} us")


if __name__ == "__main__":
    main(sys.argv[1])
