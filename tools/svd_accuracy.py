#!/usr/bin/env python
"""Accuracy and time of the three drivers of the small Jacobi SVD (option "svd.fused": 0 launch per round,
2 flat single-cluster launch, 1 block driver) against LAPACK's dgesdd on graded triangular matrices.
    python tools/svd_accuracy.py > profiles/r02/svd_drivers.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsi_b200 as gsi      # noqa: E402

ctx = gsi.default_context()
out = []
for l in (60, 110, 129, 210, 256):
    rng = np.random.default_rng(l)
    M = np.triu(rng.standard_normal((l, l))) * (10.0 ** (-6 * np.arange(l) / max(l - 1, 1)))[:, None]
    sref = np.linalg.svd(M, compute_uv=False)
    row = {"l": l}
    for mode in (0, 2, 1):
        ctx.set_option("svd.fused", mode)
        U, s = gsi.svd_small(M)
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(3):
            U, s = gsi.svd_small(M)
        ctx.sync()
        row[f"mode{mode}"] = {"ms": (time.perf_counter() - t0) / 3 * 1e3,
                             "sv_err_over_s1_eps": float(np.max(np.abs(s - sref)) / sref[0] / 2.220446049250313e-16),
                             "orth_err_eps": float(np.max(np.abs(U.T @ U - np.eye(l))) / 2.220446049250313e-16),
                             "sweeps": ctx.get_option("svd.last_sweeps")}
    ctx.set_option("svd.fused", 1)
    out.append(row)
print(json.dumps(out, indent=1))
