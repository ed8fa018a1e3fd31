#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, avg us, share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        agg.setdefault(row["Kernel Name"].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':42s} {'n':>4s} {'avg us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:42]:42s} {len(v):4d} {sum(v) / len(v):10.1f} {100 * sum(v) / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
