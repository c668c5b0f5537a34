"""Parity tests of code paths that are written but NOT yet verified on hardware (opt-in
options, default off).  They are skipped unless GSI_EXPERIMENTAL=1, so that the default
`-m gpu` run only exercises verified code:

    GSI_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q
"""
import os

import numpy as np
import pytest

import oracle
from gpu_util import gsi, relerr  # noqa: F401

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("GSI_EXPERIMENTAL") != "1",
                                 reason="experimental options: set GSI_EXPERIMENTAL=1")]


@pytest.mark.parametrize("n,l", [(64, 8), (300, 17), (1000, 60), (5000, 110), (20000, 210), (777, 256), (40, 33)])
def test_lu_fused_matches_per_column(gsi, n, l):
    """"lu.fused": the cooperative single-launch panel performs the same pivot choices and the
    same arithmetic as the launch-per-column path -> bit-identical L; and both match the
    oracle's dgetrf-based unpermuted L."""
    from gsi_b200.pcga import lu_L
    ctx = gsi.default_context()
    rng = np.random.default_rng(n + l)
    Y = rng.standard_normal((n, l))
    saved = ctx.get_option("lu.fused")
    try:
        ctx.set_option("lu.fused", 0)
        L0 = lu_L(Y)
        ctx.set_option("lu.fused", 1)
        L1 = lu_L(Y)
    finally:
        ctx.set_option("lu.fused", saved)
    assert np.array_equal(L0, L1)
    assert relerr(L1, oracle.lu_L_unpermuted(Y)) < 1e-10


def test_lu_fused_ties_and_zero_pivot(gsi):
    """LAPACK tie rule (first row of maximal |value|) and the SingularException mapping."""
    from gsi_b200.pcga import lu_L
    ctx = gsi.default_context()
    Y = np.ones((200, 6))
    Y[:, 1] = np.arange(200) % 7
    Y[:, 2] = -(np.arange(200) % 5)
    Y[:, 3:] = np.random.default_rng(0).integers(-3, 4, size=(200, 3))
    saved = ctx.get_option("lu.fused")
    try:
        ctx.set_option("lu.fused", 1)
        try:
            L1 = lu_L(Y)
            err1 = None
        except gsi.SingularException as e:
            L1, err1 = None, e
        ctx.set_option("lu.fused", 0)
        try:
            L0 = lu_L(Y)
            err0 = None
        except gsi.SingularException as e:
            L0, err0 = None, e
    finally:
        ctx.set_option("lu.fused", saved)
    assert (err0 is None) == (err1 is None)
    if err0 is None:
        assert np.array_equal(L0, L1)


def test_randsvd_with_fused_lu(gsi):
    grid, ell, K, p, q = (40, 30), [6.0, 4.0], 40, 5, 2
    coords = oracle.grid_coords(grid)
    Omega = np.random.default_rng(3).standard_normal((coords.shape[1], K + p))
    op = gsi.KernelCovMatrix("exponential", coords, ell)
    ctx = gsi.default_context()
    saved = ctx.get_option("lu.fused")
    try:
        ctx.set_option("lu.fused", 0)
        Z0 = gsi.randsvd(op, K, p, q, Omega=Omega)
        ctx.set_option("lu.fused", 1)
        Z1 = gsi.randsvd(op, K, p, q, Omega=Omega)
    finally:
        ctx.set_option("lu.fused", saved)
    assert np.array_equal(Z0, Z1)


@pytest.mark.parametrize("n,l", [(64, 8), (1000, 60), (20000, 210), (300, 256), (5000, 33)])
def test_qr_fast_house(gsi, n, l):
    """"qr.fast_house": same Householder QR with a parallel reduction of the partial dot products
    (different, still deterministic, summation order): orthonormal Q spanning range(Y), R equal
    to the default path's to rounding."""
    from gsi_b200.pcga import qr_thinQ
    ctx = gsi.default_context()
    rng = np.random.default_rng(n * 3 + l)
    Y = rng.standard_normal((n, l)) * (10.0 ** (-4 * np.arange(l) / max(l - 1, 1)))[None, :]
    saved = ctx.get_option("qr.fast_house")
    try:
        ctx.set_option("qr.fast_house", 0)
        Q0, R0 = qr_thinQ(Y, return_R=True)
        ctx.set_option("qr.fast_house", 1)
        Q1, R1 = qr_thinQ(Y, return_R=True)
        Q1b = qr_thinQ(Y)
    finally:
        ctx.set_option("qr.fast_house", saved)
    assert np.array_equal(Q1, Q1b)                                  # deterministic
    assert np.max(np.abs(Q1.T @ Q1 - np.eye(l))) < 1e-12
    assert relerr(Q1 @ R1, Y) < 1e-13
    assert relerr(R1, R0) < 1e-11 and relerr(Q1, Q0) < 1e-9


def test_kcov_paced_fetch_is_bit_identical(gsi):
    """"kcov.pace": the X tile arrives in four paced bulk copies instead of one -- only the
    arrival changes, the accumulation order does not."""
    ctx = gsi.default_context()
    grid, ell, l = (110, 109), [7.0, 5.0], 24
    n = grid[0] * grid[1]
    X = np.random.default_rng(5).standard_normal((n, l))
    op = gsi.GridKernelCovMatrix("exponential", grid, ell)
    saved = {k: ctx.get_option(k) for k in ("kcov.pace", "kcov.window", "kcov.sweep_div")}
    try:
        ctx.set_option("kcov.pace", 0)
        Y0 = op @ X
        ctx.set_option("kcov.pace", 4)
        Y1 = op @ X
        ctx.set_option("kcov.sweep_div", -1)
        ctx.set_option("kcov.window", 2)
        Y2 = op @ X
        ctx.set_option("kcov.pace", 0)
        Y3 = op @ X
    finally:
        for k, v in saved.items():
            ctx.set_option(k, v)
    assert np.array_equal(Y0, Y1)
    assert np.array_equal(Y2, Y3)
    coords = oracle.grid_coords(grid)
    rows = np.random.default_rng(6).choice(n, 100, replace=False)
    assert relerr(Y1[rows], oracle.kernel_cov_dense(0, coords, ell, rows=rows) @ X) < 1e-12


def test_kcov_l2_prefetch_is_bit_identical(gsi):
    """"kcov.prefetch": an extra L2 bulk prefetch ahead of each CTA's sweep moves no result."""
    ctx = gsi.default_context()
    grid, ell, l = (110, 109), [7.0, 5.0], 24
    n = grid[0] * grid[1]
    X = np.random.default_rng(5).standard_normal((n, l))
    op = gsi.GridKernelCovMatrix("gaussian", grid, ell)
    saved = ctx.get_option("kcov.prefetch")
    try:
        ctx.set_option("kcov.prefetch", 0)
        Y0 = op @ X
        for ahead in (1, 3, 400):              # 400 > number of k-tiles: the prefetch is skipped
            ctx.set_option("kcov.prefetch", ahead)
            assert np.array_equal(op @ X, Y0)
    finally:
        ctx.set_option("kcov.prefetch", saved)
