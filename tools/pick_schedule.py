#!/usr/bin/env python
"""Reads the sweep_probe.py result and prints the GSI_SWEEP string of the schedule to adopt:
the fastest windowed schedule if it is no more than 0.5 % slower than the unthrottled round-1
default (first entry), else that default."""
import json
import sys

rows = json.load(open(sys.argv[1]))["schedules"]
base = rows[0]
windowed = [r for r in rows if r["window"] > 0 and r["tflops"]]
best = max(windowed, key=lambda r: r["tflops"]) if windowed else base
pick = best if best["tflops"] >= 0.995 * base["tflops"] else base
print("{groups},{div},{hint},{window},{epoch_shift}".format(**pick))
