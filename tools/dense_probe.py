#!/usr/bin/env python
"""Times (and, under ncu, exposes) single launches of the dense tall-skinny DMMA GEMM
(csrc/dense_gemm.cu): A (n x n, column-major, random) times X (n x l), N and T variants.

    python tools/dense_probe.py [--n 16384] [--l 210] [--reps 3]
    ncu --set full --import-source on --clock-control none -k regex:dense_gemm --launch-skip 2 -c 2 \
        -o gpurun_out/prof_dense python tools/dense_probe.py --reps 1
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
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--l", type=int, default=210)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    ctx = gsi.default_context()
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((args.n, args.n)))
    op = gsi.DenseMatrix(A, ctx=ctx)
    X = gsi.DeviceMatrix.from_host(ctx, rng.standard_normal((args.n, args.l)))
    out = {"n": args.n, "l": args.l}
    for name, tr in (("N", False), ("T", True)):
        op.apply(X, trans=tr).free()                       # warm-up launch (skipped by the ncu command above)
        ctx.sync()
        ctx.gemm_timing(enable=True)
        for _ in range(args.reps):
            op.apply(X, trans=tr).free()
        ctx.sync()
        ms, nl, fl = ctx.gemm_timing(enable=False)
        out[name] = {"ms_per_launch": ms / max(nl, 1), "tflops": fl / (ms * 1e-3) * 1e-12 if ms > 0 else None}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
