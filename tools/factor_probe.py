#!/usr/bin/env python
"""Time of one LU (lu(Y).L) and one thin QR of a resident n x l iterate (the normalisers of the power iteration):
    python tools/factor_probe.py [n] [l]        # default 1000000 x 210: the 10^6-point case, rows beyond shared memory"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsi_b200 as gsi      # noqa: E402
from gsi_b200._lib import check  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
l = int(sys.argv[2]) if len(sys.argv) > 2 else 210
ctx = gsi.default_context()
A = gsi.DeviceMatrix.from_host(ctx, np.random.default_rng(0).standard_normal((n, l)))
B = gsi.DeviceMatrix(ctx, n, l)
out = {"n": n, "l": l}
for name, call in (("lu_ms", lambda: ctx._lib.gsi_lu_L(ctx._h, B._h)),
                   ("qr_ms", lambda: ctx._lib.gsi_qr_thinQ(ctx._h, B._h, None, l))):
    ts = []
    for _ in range(3):
        check(ctx._lib.gsi_buf_copy(A._h, B._h))
        ctx.sync()
        t0 = time.perf_counter()
        check(call())
        ctx.sync()
        ts.append((time.perf_counter() - t0) * 1e3)
    out[name] = min(ts)
print(json.dumps(out))
