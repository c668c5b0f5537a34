"""Timings of the BASELINE.json configurations that are parity cases rather than the bench
headline: C1 dense rank-50 randsvd, C2 pcgalsqr on a 100x100 exponential
covariance (randsvd + per-iteration split), C4 LowRankCovMatrix prior + rga sketch products.
GPU path through the C ABI next to the CPU oracle (SciPy/OpenBLAS, all host cores).

    python tools/bench_configs.py > profiles/r01/configs_c1_c2_c4.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsi_b200 as gsi      # noqa: E402
import oracle               # noqa: E402  (reported CPU baseline / checker only)
from oracle.fftrf import powerlaw_structuredgrid  # noqa: E402

ctx = gsi.default_context()


def best(f, reps=3, warm=1):
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


out = {"cores": os.cpu_count()}

# ---- C1: dense 1000x1000 rank-50, K=50 p=10 q=2
rng = np.random.default_rng(2017)
A = rng.standard_normal((1000, 50)) @ rng.standard_normal((50, 1000))
Om = np.random.default_rng(0).standard_normal((1000, 60))
opA = gsi.DenseMatrix(A)
Omd = gsi.DeviceMatrix.from_host(ctx, Om)
Z = gsi.randsvd(opA, 50, 10, 2, Omega=Om)
c = oracle.compare_Z(Z, oracle.randsvd(A, Om, 50, 10, 2), 50)
out["c1"] = {"gpu_ms_resident": best(lambda: gsi.randsvd(opA, 50, 10, 2, Omega=Omd, device_out=True).free(), 5, 3),
             "gpu_ms_host_arrays": best(lambda: gsi.randsvd(opA, 50, 10, 2, Omega=Om), 5, 2),
             "cpu_oracle_ms": best(lambda: oracle.randsvd(A, Om, 50, 10, 2), 3, 1), "parity": c}

# ---- C2: 100x100 grid, exponential covariance, 200 linear observations, K=100 p=10 q=3
grid, nobs, K, p, q = (100, 100), 200, 100, 10, 3
coords = oracle.grid_coords(grid)
n = coords.shape[1]
ell = [12.0, 8.0]
Om2 = np.random.default_rng(1).standard_normal((n, K + p))
opC = gsi.GridKernelCovMatrix("exponential", grid, ell)
t_gpu = best(lambda: gsi.randsvd(opC, K, p, q, Omega=Om2), 3, 2)
Cd = oracle.kernel_cov_dense(0, coords, ell)
t_cpu = best(lambda: oracle.randsvd(Cd, Om2, K, p, q), 2, 1)
xis = gsi.getxis(opC, K, p, q, Omega=Om2)
xis_ref = oracle.getxis(Cd, Om2, K, p, q)
par = max(min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b) for a, b in zip(xis, xis_ref))
opD = gsi.DenseMatrix(Cd)                       # the reference's own form: dense Q::Matrix
t_gpu_dense = best(lambda: gsi.randsvd(opD, K, p, q, Omega=Om2), 3, 2)
H = rng.standard_normal((nobs, n)) / np.sqrt(n)
mu = 2.0
truth = mu + np.stack(xis, axis=1) @ rng.standard_normal(K)
noise = 1e-4
y = H @ truth + noise * rng.standard_normal(nobs)
R = noise ** 2 * np.ones(nobs)
X = np.ones(n)
s0 = np.full(n, mu)
from gsi_b200.pcga import LinearForwardModel, pcgalsqriteration, _xis_to_device  # noqa: E402
fm = LinearForwardModel(H)
dev = _xis_to_device(ctx, xis)
delta = float(np.sqrt(np.finfo(float).eps))
t_it_gpu = best(lambda: pcgalsqriteration(fm, s0, X, xis, R, y, delta, ctx=ctx, _dev=dev), 5, 2)
t_it_gpu_host_model = best(lambda: pcgalsqriteration(lambda s: H @ s, s0, X, xis, R, y, delta, ctx=ctx, _dev=dev), 3, 1)
t_it_cpu = best(lambda: oracle.pcgalsqriteration(lambda s: H @ s, s0, X, xis, R, y, delta), 3, 1)
s_gpu = gsi.pcgalsqr(fm, s0, X, xis, R, y)
out["c2"] = {"n": n, "randsvd_gpu_ms_matrixfree": t_gpu, "randsvd_gpu_ms_dense_operator": t_gpu_dense,
             "randsvd_cpu_oracle_ms": t_cpu, "xis_parity_up_to_sign": par,
             "pcgalsqr_iteration_gpu_ms_declared_linear_model": t_it_gpu,
             "pcgalsqr_iteration_gpu_ms_host_blackbox_model": t_it_gpu_host_model,
             "pcgalsqr_iteration_cpu_oracle_ms": t_it_cpu,
             "estimate_rel_err_vs_truth": float(np.linalg.norm(s_gpu - truth) / np.linalg.norm(truth))}

# ---- C4: 256x256 power-law fields -> LowRankCovMatrix prior; rga sketch products 500 x 1e5
rng = np.random.default_rng(4)
nf = 200
fields = [powerlaw_structuredgrid([256, 256], 2.0, 3.14, -3.5, rng).ravel(order="F") for _ in range(nf)]
n4 = 256 * 256
Om4 = np.random.default_rng(5).standard_normal((n4, 110))
lr = gsi.LowRankCovMatrix(fields)
t_lr_gpu = best(lambda: gsi.randsvd(lr, 100, 10, 3, Omega=Om4), 3, 2)
lro = oracle.LowRankCovMatrix(fields)
t0 = time.perf_counter()
Zo = oracle.randsvd(lro, Om4, 100, 10, 3)
t_lr_cpu = (time.perf_counter() - t0) * 1e3
Zg = gsi.randsvd(lr, 100, 10, 3, Omega=Om4)
c4 = oracle.compare_Z(Zg, Zo, 100)
Nred, nobs4 = 500, 100000
S = rng.standard_normal((Nred, nobs4)) / np.sqrt(nobs4)
Rd = np.full(nobs4, 1e-8)
from gsi_b200.pcga import _Sketch  # noqa: E402
sk = _Sketch(S, ctx)
V = np.asfortranarray(rng.standard_normal((nobs4, 103)))     # as rga assembles the K+3 forward runs (column-major)
t_cov = best(lambda: sk.cov(Rd), 3, 1)
t_app = best(lambda: sk.apply(V), 3, 1)
t_cov_cpu = best(lambda: (S * Rd[None, :]) @ S.T, 2, 1)
t_app_cpu = best(lambda: S @ V, 2, 1)
out["c4"] = {"lowrankcov_randsvd_gpu_ms": t_lr_gpu, "lowrankcov_randsvd_cpu_oracle_ms": t_lr_cpu, "parity": c4,
             "sketch_cov_gpu_ms": t_cov, "sketch_cov_cpu_ms": t_cov_cpu,
             "sketch_apply_103cols_gpu_ms_incl_transfers": t_app, "sketch_apply_103cols_cpu_ms": t_app_cpu}
# ---- a7 / f4: adaptive range finder on a dense 8192^2 matrix of rank 96 (HBM-bound: one pass over A per vector)
na, ra = 8192, 96
rng = np.random.default_rng(7)
Aa = rng.standard_normal((na, ra)) @ rng.standard_normal((ra, na))
opa = gsi.DenseMatrix(Aa)
Om0, oms = rng.standard_normal((na, 10)), rng.standard_normal((na, 160))
Qf = gsi.rangefinder(opa, Omega=Om0, omegas=oms)
t_f = best(lambda: gsi.rangefinder(opa, Omega=Om0, omegas=oms), 3, 1)
passes = Qf.shape[1] + 1                      # A * randn(n, r) once, then A * omega per basis vector
Qb = gsi.rangefinder(opa, omegas=oms, block=32)
t_b = best(lambda: gsi.rangefinder(opa, omegas=oms, block=32), 3, 1)
t_o = best(lambda: oracle.rangefinder_adaptive(Aa, Om0, oms), 1, 0)
out["adaptive"] = {"n": na, "rank": ra,
                   "faithful_ms": t_f, "faithful_basis": int(Qf.shape[1]), "faithful_passes_over_A": passes,
                   "faithful_algorithmic_GBps": passes * 8.0 * na * na / (t_f * 1e-3) * 1e-9,
                   "faithful_residual": float(np.linalg.norm(Aa - Qf @ (Qf.T @ Aa)) / np.linalg.norm(Aa)),
                   "blocked32_ms": t_b, "blocked32_basis": int(Qb.shape[1]),
                   "blocked32_residual": float(np.linalg.norm(Aa - Qb @ (Qb.T @ Aa)) / np.linalg.norm(Aa)),
                   "cpu_oracle_faithful_ms": t_o}
print(json.dumps(out, indent=1))
