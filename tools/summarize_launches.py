#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv`):
per-kernel time, DRAM bytes and share of the LAST complete MSM call (the launches between two k_normalize).

    tools/summarize_launches.py profiles/r02_launches_n24.csv [--traffic-json profiles/traffic.json --logn 24]

With --traffic-json the summed DRAM traffic of the bucket-accumulation phase (affine levels + work list + k_accumulate,
what bench.py's roofline.traffic reports) is merged into that file."""
import argparse
import csv
import json
import os
from collections import OrderedDict

ACC_PHASE = ("k_aff_prepare", "k_aff_invert", "k_aff_finish", "k_classify", "k_size_scan", "k_worklist_fill", "k_accumulate")
UNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def load(path):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rows = list(csv.reader(lines))
    ix = {h: i for i, h in enumerate(rows[0])}
    launches = OrderedDict()
    for r in rows[1:]:
        key = int(r[ix["ID"]])
        d = launches.setdefault(key, {"name": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")})
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            d["us"] = val * UNIT.get(unit, 1.0)
        elif m.startswith("dram__bytes"):
            d[m] = val * BYTES.get(unit, 1.0)
        else:
            d[m] = val
    return list(launches.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic-json")
    ap.add_argument("--logn", type=int)
    a = ap.parse_args()
    rows = load(a.csv)
    idx = [i for i, r in enumerate(rows) if "k_normalize" in r["name"]]
    print(f"{a.csv}: {len(rows)} launches, {len(idx)} MSM calls")
    if len(idx) < 2:
        return
    seg = rows[idx[-2] + 1: idx[-1] + 1]
    tot = sum(r.get("us", 0) for r in seg)
    phase_us = phase_bytes = 0.0
    for r in seg:
        rd, wr = r.get("dram__bytes_read.sum", 0), r.get("dram__bytes_write.sum", 0)
        in_phase = any(k in r["name"] for k in ACC_PHASE)
        if in_phase:
            phase_us += r.get("us", 0)
            phase_bytes += rd + wr
        print(f"  {r['name']:<28} {r.get('us', 0):10.1f} us {100 * r.get('us', 0) / tot:5.1f} %   rd {rd / 1e9:7.3f} GB  wr {wr / 1e9:7.3f} GB  "
              f"sm {r.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f} %{'  *' if in_phase else ''}")
    print(f"  {'total':<28} {tot:10.1f} us;  bucket-accumulation phase (*): {phase_us:.1f} us = {100 * phase_us / tot:.1f} % of the call, "
          f"{phase_bytes / 1e9:.2f} GB of DRAM traffic")
    if a.traffic_json and a.logn:
        data = {}
        if os.path.exists(a.traffic_json):
            with open(a.traffic_json) as fh:
                data = json.load(fh)
        data.setdefault("phase_traffic_bytes", {})[str(a.logn)] = phase_bytes
        data["source"] = "profiles/r02_launches_n{logn}.csv (ncu launch list of `bench.py --logn {logn}`, summed by tools/summarize_launches.py)"
        with open(a.traffic_json, "w") as fh:
            json.dump(data, fh, indent=1)


if __name__ == "__main__":
    main()
