#!/usr/bin/env python
"""Times single launches of the matrix-free product kernel (csrc/kcov_gemm.cu) on a structured grid.

    python tools/kcov_probe.py --grid 64,56,56 [--generation table|arithmetic] [--reps 3] [--l 210]
    python tools/kcov_probe.py --grid 32,28,28      # 25088 rows = the row schedule of one rank of C3 at 8 GPUs
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsi_b200 as gsi      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="64,56,56")
    ap.add_argument("--ell", default="9,7,5")
    ap.add_argument("--kind", default="gaussian")
    ap.add_argument("--generation", default="table", choices=["table", "arithmetic"])
    ap.add_argument("--l", type=int, default=210)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    grid = [int(v) for v in args.grid.split(",")]
    ell = [float(v) for v in args.ell.split(",")][:len(grid)]
    n = int(np.prod(grid))
    ctx = gsi.default_context()
    if args.generation == "table":
        op = gsi.GridKernelCovMatrix(args.kind, grid, ell, ctx=ctx)
    else:
        axes = np.meshgrid(*[np.arange(s, dtype=np.float64) for s in grid], indexing="ij")
        coords = np.stack([g.ravel(order="F") for g in axes], axis=0)
        op = gsi.KernelCovMatrix(args.kind, coords, ell, ctx=ctx)
    X = gsi.DeviceMatrix.from_host(ctx, np.random.default_rng(0).standard_normal((n, args.l)))
    op.apply(X).free()
    ctx.sync()
    ctx.gemm_timing(enable=True)
    for _ in range(args.reps):
        op.apply(X).free()
    ctx.sync()
    ms, nl, fl = ctx.gemm_timing(enable=False)
    rg = (n + 15) // 16
    print(json.dumps({"grid": grid, "n": n, "l": args.l, "generation": args.generation, "row_groups": rg,
                      "rounds_of_592": rg / 592.0, "ms_per_launch": ms / max(nl, 1),
                      "tflops": fl / (ms * 1e-3) * 1e-12 if ms > 0 else None}))


if __name__ == "__main__":
    main()
