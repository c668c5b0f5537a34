#!/usr/bin/env python
"""Digest of the source page of an `ncu --set full --import-source on` capture, summed over all captured launches
of a kernel: issue-stall shares, warp-instructions executed, and the SASS instructions that collect the most
samples (with their dominant stall reasons).

    ncu -i capture.ncu-rep --page source --csv > capture_source.csv
    python tools/ncu_source_digest.py capture_source.csv [top_n] > profiles/rNN/<kernel>_source_digest.txt"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    tables, cur, names = [], None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            tables.append(cur)
            names.append(r[1] if len(r) > 1 else "")
            continue
        if cur is not None:
            cur.append(r)
    hdr = tables[0][0]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    n = len([r for r in tables[0][1:] if len(r) == len(hdr)])
    samp, ex, src = [0] * n, [0] * n, [None] * n
    st = [collections.Counter() for _ in range(n)]
    used = 0
    for t in tables:
        body = [r for r in t[1:] if len(r) == len(hdr)]
        if len(body) != n:
            continue
        used += 1
        for i, r in enumerate(body):
            samp[i] += int(r[idx["# Samples"]])
            ex[i] += int(r[idx["Instructions Executed"]])
            src[i] = r[idx["Source"]].strip()
            for h in stalls:
                v = int(r[idx[h]] or 0)
                if v:
                    st[i][h[6:]] += v
    total = sum(samp)
    print(f"kernel: {names[0]}")
    print(f"launches summed: {used}   SASS instructions: {n}   samples: {total}   warp-instructions executed: {sum(ex)}")
    agg = collections.Counter()
    for c in st:
        agg.update(c)
    tot = sum(agg.values())
    print("issue-stall shares (all samples):")
    for k, v in agg.most_common(10):
        print(f"  {k:22s} {100.0 * v / tot:5.1f} %")
    print(f"instructions with the most samples (top {top_n}, program order):")
    keep = sorted(sorted(range(n), key=lambda i: -samp[i])[:top_n])
    for i in keep:
        dom = ", ".join(f"{k} {v}" for k, v in st[i].most_common(2))
        print(f"  #{i:5d} {100.0 * samp[i] / total:5.1f} %  executed {ex[i]:9d}  {src[i][:64]:64s} [{dom}]")


if __name__ == "__main__":
    main()
