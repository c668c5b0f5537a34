#!/usr/bin/env python
"""Latency-bound BASELINE configurations, one at a time (the target of an ncu launch list):

    python tools/small_probe.py c1|c2|c2dense|sketch [reps]

c1       dense 1000x1000 rank-50, K=50 p=10 q=2 (BASELINE configs[0]), Omega resident
c2       100x100 exponential covariance, K=100 p=10 q=3, matrix-free (configs[1]'s prior)
c2dense  the same with the covariance as a dense A::Matrix (the reference's own form)
sketch   rga sketch product S(500 x 1e5) * V(1e5 x 103): pageable source, page-locked source
Prints one JSON line with the best-of-reps host time (ms) and the phase times the library records.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsi_b200 as gsi      # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = gsi.default_context()


def best(f, reps=reps, warm=2):
    for _ in range(warm):
        f()
    ctx.sync()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def phases(f):
    ctx.gemm_timing(enable=True)
    ctx.phase_timing(reset=True)
    ctx.launch_count(reset=True)
    f()
    ctx.sync()
    ctx.gemm_timing(enable=False)
    ph = ctx.phase_timing()
    ph["launches"] = ctx.launch_count()
    ph["svd_sweeps"] = ctx.get_option("svd.last_sweeps")
    return ph


out = {"config": which}
if which == "c1":
    rng = np.random.default_rng(2017)
    A = rng.standard_normal((1000, 50)) @ rng.standard_normal((50, 1000))
    op = gsi.DenseMatrix(A)
    Omd = gsi.DeviceMatrix.from_host(ctx, np.random.default_rng(0).standard_normal((1000, 60)))
    f = lambda: gsi.randsvd(op, 50, 10, 2, Omega=Omd, device_out=True).free()      # noqa: E731
    out["ms"] = best(f)
    out["phases"] = phases(f)
elif which in ("c2", "c2dense"):
    grid, K, p, q = (100, 100), 100, 10, 3
    ell = [12.0, 8.0]
    n = grid[0] * grid[1]
    if which == "c2":
        op = gsi.GridKernelCovMatrix("exponential", grid, ell)
    else:
        ax = [np.arange(s, dtype=np.float64) for s in grid]
        gx, gy = np.meshgrid(*ax, indexing="ij")
        c = np.stack([gx.ravel(order="F") / ell[0], gy.ravel(order="F") / ell[1]], axis=0)
        d2 = (c[0][:, None] - c[0][None, :]) ** 2 + (c[1][:, None] - c[1][None, :]) ** 2
        op = gsi.DenseMatrix(np.asfortranarray(np.exp(-np.sqrt(d2))))
    Omd = gsi.DeviceMatrix.from_host(ctx, np.random.default_rng(1).standard_normal((n, K + p)))
    f = lambda: gsi.randsvd(op, K, p, q, Omega=Omd, device_out=True).free()        # noqa: E731
    out["ms"] = best(f)
    out["phases"] = phases(f)
elif which == "sketch":
    from gsi_b200.pcga import _Sketch
    rng = np.random.default_rng(4)
    Nred, nobs, c = 500, 100000, 103
    S = rng.standard_normal((Nred, nobs)) / np.sqrt(nobs)
    sk = _Sketch(S, ctx)
    V = np.asfortranarray(rng.standard_normal((nobs, c)))
    Vp = sk.batch_buffer(nobs, c)
    Vp[...] = V
    r0 = sk.apply(V)
    r1 = sk.apply(Vp)
    out["pageable_ms"] = best(lambda: sk.apply(V))
    out["pinned_ms"] = best(lambda: sk.apply(Vp))
    t0 = time.perf_counter()
    ref = S @ V
    out["cpu_gemm_ms"] = (time.perf_counter() - t0) * 1e3
    out["rel_err_pageable"] = float(np.abs(r0 - ref).max() / np.abs(ref).max())
    out["rel_err_pinned"] = float(np.abs(r1 - ref).max() / np.abs(ref).max())
print(json.dumps(out))
