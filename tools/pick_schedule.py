#!/usr/bin/env python
"""Reads a sweep_probe.py result.  Prints two lines:
  1. the GSI_SWEEP string of the schedule to adopt: among the windowed schedules that are no more
     than 0.5 % slower than the unthrottled round-1 default (first entry), the one expected to move
     the fewest DRAM bytes (a single front, else the fewest fronts); the default if none qualifies;
  2. the schedule-set indices worth a DRAM-bytes capture (default, the pick, the next two fastest).
"""
import json
import sys

rows = json.load(open(sys.argv[1]))["schedules"]
base = rows[0]
ok = [r for r in rows if r["window"] > 0 and r["tflops"] and r["tflops"] >= 0.995 * base["tflops"]]


def fronts(r):
    return 1 if r["div"] <= 0 else r["groups"]


pick = min(ok, key=lambda r: (fronts(r), -r["tflops"])) if ok else base
fast = sorted((r for r in rows if r["window"] > 0 and r["tflops"]), key=lambda r: -r["tflops"])
idx = []
for r in [base, pick] + fast:
    if r["index"] not in idx:
        idx.append(r["index"])
print("{groups},{div},{hint},{window},{epoch_shift}".format(**pick))
print(" ".join(str(i) for i in idx[:4]))
