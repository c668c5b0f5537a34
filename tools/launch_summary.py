#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py gpurun_out/launches.csv [steps_in_list]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        k = d["Kernel Name"][:64]
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{len(data)} launches, {tot / div:.3f} ms of kernel time per step (list holds {div:g} steps)")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:66s} {n / div:8.1f} launches {t / div:10.3f} ms {100 * t / tot:6.2f}%  {1e3 * t / n:9.1f} us each")


if __name__ == "__main__":
    main()
