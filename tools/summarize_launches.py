#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share of the
last complete MSM call (the launches between two k_normalize).  Usage: summarize_launches.py file.csv"""
import csv
import sys


def main(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        try:
            rows.append((r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
        except (KeyError, ValueError):
            pass
    idx = [i for i, r in enumerate(rows) if "k_normalize" in r[0]]
    print(f"{path}: {len(rows)} launches, {len(idx)} MSM calls")
    if len(idx) < 2:
        return
    seg = rows[idx[-2] + 1: idx[-1] + 1]
    scale = 1e-3 if seg[0][2] in ("ns", "nsecond") else 1.0
    tot = sum(r[1] for r in seg)
    for name, v, _ in seg:
        print(f"  {name:<28} {v * scale:12.1f} us {100 * v / tot:6.1f} %")
    print(f"  {'total':<28} {tot * scale:12.1f} us")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
