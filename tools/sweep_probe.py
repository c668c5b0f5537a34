#!/usr/bin/env python
"""Times ONE launch shape of the matrix-free product kernel (kcov_gemm_kernel) under different
k-sweep schedules (gsi_ctx_set_option "kcov.*"), in one process:

    python tools/sweep_probe.py [--workload c3] [--reps 2] [--out gpurun_out/sweep_probe.json]
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:kcov \
        --csv --log-file gpurun_out/sweep_probe_ncu.csv python tools/sweep_probe.py --reps 1 --no-warm

The launch order printed in the JSON is the launch order ncu sees, so the DRAM bytes of each
schedule can be read off the ncu CSV by position.  Each entry: groups, div, hint, window,
epoch_shift -> ms per launch (CUDA events on the library stream) and TFLOP/s.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402

# (sweep_groups, sweep_div, l2_hint, window, epoch_shift)
SCHEDULE_SETS = {
    # first pass (profiles/r01/sweep_schedules_c3_summary.csv): ONE coherent front (all CTAs within a few
    # MB of X) cuts DRAM traffic 611 -> 8 GB per launch but runs 13 % slower, whatever the window
    "coherent": [
        (64, 256, 0, 0, 6),      # round-1 default: de-synchronised over a quarter of X, unthrottled
        (64, -1, 0, 4, 6),       # 1-tile separation, window of 4 x 64 tiles
        (64, -1, 0, 16, 6),
        (64, -1, 0, 2, 5),
        (1, 0, 0, 4, 6),         # lock-step starts
        (64, -4, 0, 4, 6),       # 4-tile separation
        (64, -1, 1, 8, 6),       # + evict_last hint
        (64, -1, 0, 64, 6),      # loose window (64 x 64 tiles = 240 MB: only stops runaway drift)
        (64, -1, 0, 2, 4),       # tight: 2 x 16 tiles
        (1, 0, 0, 16, 6),        # lock-step starts, wider window
        (16, -1, 0, 8, 6),       # 16 start offsets
        (64, 256, 0, 4, 6),      # quarter-of-X spread but bounded drift
    ],
    # second pass: G fronts spread evenly over X (each front L2-resident: DRAM ~ G x 8 GB), and
    # single fronts whose CTAs are de-phased by a few tiles each
    "fronts": [
        (64, 256, 0, 0, 6),
        (8, 8, 0, 2, 5),
        (16, 16, 0, 2, 4),
        (32, 32, 0, 1, 3),
        (4, 4, 0, 2, 6),
        (2, 2, 0, 4, 6),
        (64, -8, 0, 2, 6),       # one front, 64 positions 8 tiles apart
        (256, -3, 0, 2, 6),      # one front, every CTA at its own position 3 tiles apart
        (16, 16, 0, 1, 4),
        (8, 8, 0, 4, 4),
    ],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--no-warm", action="store_true")
    ap.add_argument("--control", action="store_true", help="also time an L2-resident problem (X = 50 MB)")
    ap.add_argument("--generation", default="table", choices=["table", "arithmetic"],
                    help="kernel values from the lattice table (default) or from coordinates (no table look-ups "
                         "on the L2 return path: separates the two suspects of the L2-served-stream penalty)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_probe.json"))
    ap.add_argument("--set", default="coherent", choices=sorted(SCHEDULE_SETS))
    ap.add_argument("--only", type=int, nargs="*", default=None, help="indices into the schedule set")
    args = ap.parse_args()

    import gsi_b200 as gsi
    kind, grid, ell, K, p, q, desc = WORKLOADS[args.workload]
    l = K + p
    n = int(np.prod(grid))
    ctx = gsi.default_context()
    if args.generation == "table":
        op = gsi.GridKernelCovMatrix(kind, grid, ell, ctx=ctx)
    else:
        from bench import grid_coords
        op = gsi.KernelCovMatrix(kind, grid_coords(grid), ell, ctx=ctx)
    X = gsi.DeviceMatrix.from_host(ctx, np.random.default_rng(0).standard_normal((n, l)))
    Y = gsi.DeviceMatrix(ctx, n, l)
    lib = ctx._lib

    def apply():
        gsi._lib.check(lib.gsi_op_apply(op._h, 0, X._h, Y._h))

    rows = []
    ref = None
    for i, (g, d, h, w, es) in enumerate(SCHEDULE_SETS[args.set]):
        if args.only is not None and i not in args.only:
            continue
        ctx.set_option("kcov.window", 0)
        ctx.set_option("kcov.sweep_groups", g)
        ctx.set_option("kcov.sweep_div", d)
        ctx.set_option("kcov.l2_hint", h)
        ctx.set_option("kcov.epoch_shift", es)
        ctx.set_option("kcov.window", w)
        if not args.no_warm:
            apply()
        ctx.sync()
        ctx.gemm_timing(enable=True)
        for _ in range(args.reps):
            apply()
        ms, nl, fl = ctx.gemm_timing(enable=False)
        # spot check: a few rows must agree between schedules to rounding (the k order of a CTA
        # depends on its sweep start, so only schedules with equal (groups, div) are bit-identical)
        y = Y.rows_numpy(12345, 4)
        if ref is None:
            ref = y
        dev = float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))
        rows.append({"index": i, "groups": g, "div": d, "hint": h, "window": w, "epoch_shift": es,
                     "launches": nl + (0 if args.no_warm else 1), "ms_per_launch": ms / max(nl, 1),
                     "tflops": fl / (ms * 1e-3) * 1e-12 if ms > 0 else None, "rel_dev_vs_first": dev})
        print(json.dumps(rows[-1]), flush=True)
    control = None
    if args.control:
        # Control: the same kernel shape on a problem whose X (50 MB) stays L2-resident whatever the
        # schedule -- is an X stream served entirely from L2 as fast per k-tile as one served from HBM?
        X.free(); Y.free(); op.free()
        cgrid = (32, 32, 28)
        cn = int(np.prod(cgrid))
        cop = gsi.GridKernelCovMatrix(kind, cgrid, ell, ctx=ctx)
        cX = gsi.DeviceMatrix.from_host(ctx, np.random.default_rng(1).standard_normal((cn, l)))
        cY = gsi.DeviceMatrix(ctx, cn, l)
        control = []
        for (g, d, h, w, es) in [(64, 256, 0, 0, 6), (64, -1, 0, 4, 6), (1, 0, 0, 0, 6)]:
            ctx.set_option("kcov.window", 0)
            ctx.set_option("kcov.sweep_groups", g)
            ctx.set_option("kcov.sweep_div", d)
            ctx.set_option("kcov.epoch_shift", es)
            ctx.set_option("kcov.window", w)
            for _ in range(3):
                gsi._lib.check(lib.gsi_op_apply(cop._h, 0, cX._h, cY._h))
            ctx.sync()
            ctx.gemm_timing(enable=True)
            for _ in range(20):
                gsi._lib.check(lib.gsi_op_apply(cop._h, 0, cX._h, cY._h))
            ms, nl, fl = ctx.gemm_timing(enable=False)
            control.append({"n": cn, "groups": g, "div": d, "window": w, "epoch_shift": es,
                            "ms_per_launch": ms / nl, "tflops": fl / (ms * 1e-3) * 1e-12})
            print("control", json.dumps(control[-1]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"workload": desc, "n": n, "l": l, "generation": args.generation, "pace": args.pace, "prefetch": args.prefetch, "schedules": rows,
                   "l2fit_control": control}, f, indent=1)


if __name__ == "__main__":
    main()
