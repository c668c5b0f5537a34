"""Run under torchrun on N GPUs: row-sharded randsvd must agree with the CPU oracle and
be independent of N.  Prints one JSON line on rank 0; exit code != 0 on
failure.  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_randsvd_check.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gsi_b200 as gsi            # noqa: E402
from gsi_b200 import dist as gdist  # noqa: E402
import oracle                     # noqa: E402  (checker)


def main():
    ctx = gdist.context_from_torch_distributed()
    import torch.distributed as dist
    results = {}
    ok = True
    for name, kind, grid, ell, K, p, q in [("exp2d", "exponential", (90, 80), (12.0, 8.0), 200, 10, 2),
                                           ("gauss3d", "gaussian", (20, 18, 16), (9.0, 7.0, 5.0), 100, 10, 2),
                                           ("exp2d_q0", "exponential", (50, 40), (12.0, 8.0), 40, 8, 0)]:
        coords = oracle.grid_coords(grid)
        n = coords.shape[1]
        r0, ml = gsi.partition_rows(n, ctx.world, ctx.rank)
        op = gsi.KernelCovMatrix(kind, coords, ell, ctx=ctx, row0=r0, mloc=ml)
        Omega = np.random.default_rng(0).standard_normal((n, K + p))
        if name == "gauss3d":       # structured-grid (lattice table) operator, sharded
            op = gsi.GridKernelCovMatrix(kind, grid, ell, ctx=ctx, row0=r0, mloc=ml)
        Zfull = gsi.randsvd(op, K, p, q, Omega=Omega, full=True)
        Zloc = gsi.randsvd(op, K, p, q, Omega=Omega)
        assert Zloc.shape == (ml, K + p)
        same = bool(np.array_equal(Zloc, Zfull[r0:r0 + ml]))
        if ctx.rank == 0:
            kid = {"exponential": 0, "gaussian": 1}[kind]
            Zref = oracle.randsvd(oracle.kernel_cov_dense(kid, coords, ell), Omega, K, p, q)
            c = oracle.compare_Z(Zfull, Zref, K)
            c["local_block_equals_full"] = same
            results[name] = c
            ok = ok and c["tail_zero"] and c["sv_rel"] < 1e-10 and c["sine"] < 1e-8 and same
    # dense row-sharded operator (all-reduce path of A'Q)
    rng = np.random.default_rng(3)
    A = rng.standard_normal((900, 40)) @ rng.standard_normal((40, 700)) + 1e-3 * rng.standard_normal((900, 700))
    r0, ml = gsi.partition_rows(900, ctx.world, ctx.rank)
    opd = gsi.DenseMatrix(A[r0:r0 + ml], ctx=ctx, row0=r0, m_global=900)
    Om = rng.standard_normal((700, 38))
    Zd = gsi.randsvd(opd, 30, 8, 2, Omega=Om, full=True)
    if ctx.rank == 0:
        c = oracle.compare_Z(Zd, oracle.randsvd(A, Om, 30, 8, 2), 30)
        results["dense_sharded"] = c
        ok = ok and c["sv_rel"] < 1e-10 and c["sine"] < 1e-8
    # row-sharded LowRankCovMatrix (one N x l all-reduce per product) vs the oracle's operator
    fields = [rng.standard_normal(1500) * (1.0 + np.arange(1500) % 7) for _ in range(40)]
    r0, ml = gsi.partition_rows(1500, ctx.world, ctx.rank)
    Sall = np.stack(fields, axis=1)
    lr = gsi.LowRankCovMatrix(Sall[r0:r0 + ml], ctx=ctx, row0=r0, n_global=1500)
    Oml = rng.standard_normal((1500, 24))
    Zl = gsi.randsvd(lr, 20, 4, 2, Omega=Oml, full=True)
    if ctx.rank == 0:
        c = oracle.compare_Z(Zl, oracle.randsvd(oracle.LowRankCovMatrix(fields), Oml, 20, 4, 2), 20)
        results["lowrankcov_sharded"] = c
        ok = ok and c["sv_rel"] < 1e-10 and c["sine"] < 1e-8
    flag = [ok]
    dist.broadcast_object_list(flag, 0)
    if ctx.rank == 0:
        print(json.dumps({"world": ctx.world, "ok": ok, "results": results}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
