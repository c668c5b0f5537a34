"""GPU parity of rangefinder / randsvd against the oracle on identical (A, Omega).
Tolerances are the north-star ones: singular values 1e-10 relative, subspace sine 1e-8."""
import numpy as np
import pytest

import oracle
from gpu_util import gsi, relerr  # noqa: F401

pytestmark = pytest.mark.gpu

SV_TOL = 1e-10
SINE_TOL = 1e-8


def makeA(rng, n, m):
    return rng.standard_normal((n, m)) @ rng.standard_normal((m, n))


@pytest.mark.parametrize("n,m", [(10, 2), (10, 5), (100, 5), (100, 10), (100, 25)])
def test_rangefinder_fixed_reference_property(gsi, n, m):
    # test/testrmf.jl:16-18
    rng = np.random.default_rng(1000 * n + m)
    A = makeA(rng, n, m)
    Q = gsi.rangefinder(A, m, 2, rng=rng)
    assert abs(Q.shape[1] - m) <= 1
    assert np.linalg.norm(A - Q @ Q.T @ A) < 1e-8


def test_negative_iterations(gsi):
    rng = np.random.default_rng(0)
    A = makeA(rng, 10, 2)
    with pytest.raises(ValueError, match="numiterations should be positive, but numiterations=-1"):
        gsi.rangefinder(A, 2, -1)
    with pytest.raises(ValueError, match="numiterations should be positive"):
        gsi.randsvd(A, 2, 0, -3)


def test_config1_dense_rank50(gsi):
    """BASELINE config 1: dense 1000x1000 rank-50, K=50 p=10 q=2."""
    rng = np.random.default_rng(2017)
    A = rng.standard_normal((1000, 50)) @ rng.standard_normal((50, 1000))
    Omega = np.random.default_rng(0).standard_normal((1000, 60))
    Zref = oracle.randsvd(A, Omega, 50, 10, 2)
    Z = gsi.randsvd(A, 50, 10, 2, Omega=Omega)
    c = oracle.compare_Z(Z, Zref, 50)
    assert c["tail_zero"]
    assert c["sv_rel"] < SV_TOL and c["sine"] < SINE_TOL, c


@pytest.mark.parametrize("q", [0, 1, 2, 3])
def test_dense_covariance_full_rank_parity(gsi, q):
    """rank(A) > K+p: only a pivot-faithful LU normaliser matches the reference (F1)."""
    coords = oracle.grid_coords((40, 40))
    A = oracle.kernel_cov_dense(0, coords, [12.0, 8.0])
    K, p = 50, 10
    Omega = np.random.default_rng(q).standard_normal((1600, K + p))
    Zref = oracle.randsvd(A, Omega, K, p, q)
    Z = gsi.randsvd(A, K, p, q, Omega=Omega)
    c = oracle.compare_Z(Z, Zref, K)
    assert c["tail_zero"]
    assert c["sv_rel"] < SV_TOL and c["sine"] < SINE_TOL, c


def test_nonsquare_dense(gsi):
    rng = np.random.default_rng(4)
    A = rng.standard_normal((900, 40)) @ rng.standard_normal((40, 700)) + 1e-3 * rng.standard_normal((900, 700))
    K, p = 30, 8
    Omega = rng.standard_normal((700, K + p))
    Zref = oracle.randsvd(A, Omega, K, p, 2)
    Z = gsi.randsvd(A, K, p, 2, Omega=Omega)
    c = oracle.compare_Z(Z, Zref, K)
    assert c["sv_rel"] < SV_TOL and c["sine"] < SINE_TOL, c


@pytest.mark.parametrize("kind,grid,ell,K", [
    ("gaussian", (28, 26, 24), (9.0, 7.0, 5.0), 200),      # config 3 at reduced n (17 472 points)
    ("exponential", (90, 80), (12.0, 8.0), 200),           # config 5 at reduced n (7 200 points)
    ("powerlaw", (50, 45), (6.0, 5.0), 60),
])
def test_kernelcov_randsvd_parity(gsi, kind, grid, ell, K):
    p, q = 10, 2
    coords = oracle.grid_coords(grid)
    n = coords.shape[1]
    kid = {"exponential": 0, "gaussian": 1, "powerlaw": 2}[kind]
    C = oracle.kernel_cov_dense(kid, coords, ell)
    Omega = np.random.default_rng(0).standard_normal((n, K + p))
    Zref = oracle.randsvd(C, Omega, K, p, q)
    Z, S = gsi.randsvd(gsi.KernelCovMatrix(kind, coords, ell), K, p, q, Omega=Omega, return_singular_values=True)
    c = oracle.compare_Z(Z, Zref, K)
    assert c["tail_zero"]
    assert c["sv_rel"] < SV_TOL and c["sine"] < SINE_TOL, c
    assert np.max(np.abs(S[:K] - oracle.singvals_from_Z(Zref, K)) / S[:K]) < SV_TOL


@pytest.mark.parametrize("kind,grid,ell,K", [
    ("gaussian", (28, 26, 24), (9.0, 7.0, 5.0), 200),
    ("exponential", (90, 80), (12.0, 8.0), 200),
])
def test_grid_kernelcov_randsvd_parity(gsi, kind, grid, ell, K):
    """Same parity contract through the structured-grid (lattice table) operator."""
    p, q = 10, 2
    coords = oracle.grid_coords(grid)
    n = coords.shape[1]
    kid = {"exponential": 0, "gaussian": 1}[kind]
    C = oracle.kernel_cov_dense(kid, coords, ell)
    Omega = np.random.default_rng(0).standard_normal((n, K + p))
    Zref = oracle.randsvd(C, Omega, K, p, q)
    Z = gsi.randsvd(gsi.GridKernelCovMatrix(kind, grid, ell), K, p, q, Omega=Omega)
    c = oracle.compare_Z(Z, Zref, K)
    assert c["tail_zero"] and c["sv_rel"] < SV_TOL and c["sine"] < SINE_TOL, c


def test_qr_normaliser_is_not_reference(gsi):
    """Documented behaviour: NORMALISER_QR is the textbook iteration, different from the
    reference's when rank(A) > K+p (F1) -- but still a valid range finder."""
    coords = oracle.grid_coords((30, 30))
    A = oracle.kernel_cov_dense(0, coords, [12.0, 8.0])
    K, p = 30, 5
    Omega = np.random.default_rng(1).standard_normal((900, K + p))
    Zref = oracle.randsvd(A, Omega, K, p, 2)
    Zqr = gsi.randsvd(A, K, p, 2, Omega=Omega, normaliser=gsi.NORMALISER_QR)
    c = oracle.compare_Z(Zqr, Zref, K)
    sv = np.linalg.svd(A, compute_uv=False)[:K]
    err = np.abs(oracle.singvals_from_Z(Zqr, K) - sv) / sv
    assert np.max(err[:10]) < 1e-3 and np.max(err) < 0.5
    assert c["sine"] > 1e-6      # genuinely different subspace


def test_getxis_lowrank_vs_dense(gsi):
    """testrpcga.jl:83-102: getxis on the operator vs its dense materialisation, up to sign 1e-6."""
    from oracle.fftrf import powerlaw_structuredgrid
    rng = np.random.default_rng(0)
    fields = [powerlaw_structuredgrid([25, 25], 2.0, 3.14, -3.5, rng).ravel(order="F") for _ in range(100)]
    lrcm = gsi.LowRankCovMatrix(fields)
    full = np.eye(625) @ lrcm
    Omega = np.random.default_rng(0).standard_normal((625, 50))
    lrxis = gsi.getxis(lrcm, 30, 20, 3, Omega=Omega)
    fullxis = gsi.getxis(full, 30, 20, 3, Omega=Omega)
    refxis = oracle.getxis(oracle.LowRankCovMatrix(fields), Omega, 30, 20, 3)
    for a, b, r in zip(fullxis, lrxis, refxis):
        assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 1e-6
        assert min(np.linalg.norm(r - b), np.linalg.norm(r + b)) < 1e-6
