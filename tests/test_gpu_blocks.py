"""GPU parity tests of the building blocks, through the C ABI (ctypes), against the
CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

import oracle
from oracle.gepp_ref import gepp_L_unpermuted
from gpu_util import gsi, relerr  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (64, 8), (1000, 60), (4097, 210), (333, 256)])
def test_tall_roundtrip(gsi, shape):
    rng = np.random.default_rng(1)
    a = rng.standard_normal(shape)
    d = gsi.DeviceMatrix.from_host(gsi.default_context(), a)
    assert np.array_equal(d.numpy(), a)                 # bit exact
    assert np.array_equal(d.rows_numpy(shape[0] // 2, shape[0] - shape[0] // 2), a[shape[0] // 2:])


@pytest.mark.parametrize("shape", [(5, 3), (1001, 77), (64, 64)])
def test_colmajor_roundtrip(gsi, shape):
    rng = np.random.default_rng(2)
    a = rng.standard_normal(shape)
    d = gsi.DeviceMatrix.from_host(gsi.default_context(), a, gsi.LAYOUT_COLMAJOR)
    assert np.array_equal(d.numpy(), a)


@pytest.mark.parametrize("m,n,l", [(64, 64, 8), (1000, 1000, 60), (513, 777, 210), (100, 37, 5), (2000, 300, 256),
                                   (65, 1030, 110)])
def test_dense_apply(gsi, m, n, l):
    """`A*X` and `A'*X` (RandMatFact.jl:55,67,70,85) -- DMMA/TMA kernel vs dgemm.
    Tolerance 1e-13 relative (summation order differs)."""
    rng = np.random.default_rng(m + n + l)
    A = rng.standard_normal((m, n))
    X = rng.standard_normal((n, l))
    Xt = rng.standard_normal((m, l))
    op = gsi.DenseMatrix(A)
    assert op.shape == (m, n) and op.size(1) == m and op.size(2) == n
    assert relerr(op @ X, A @ X) < 1e-13
    assert relerr(op.T @ Xt, A.T @ Xt) < 1e-13
    assert relerr(Xt.T @ op, Xt.T @ A) < 1e-13          # Adjoint * A
    v = rng.standard_normal(n)
    assert relerr(op @ v, A @ v) < 1e-13


@pytest.mark.parametrize("kind", ["exponential", "gaussian", "powerlaw"])
@pytest.mark.parametrize("grid,l", [((40, 30), 60), ((14, 12, 10), 210), ((1000,), 8), ((33, 31), 5)])
def test_kernelcov_apply(gsi, kind, grid, l):
    """Matrix-free C*X vs the dense oracle materialisation (libm exp/sqrt) -- 1e-12."""
    rng = np.random.default_rng(len(grid) * 100 + l)
    coords = oracle.grid_coords(grid)
    d, n = coords.shape
    ell = [3.1, 2.7, 2.3][:d]
    kid = {"exponential": 0, "gaussian": 1, "powerlaw": 2}[kind]
    C = oracle.kernel_cov_dense(kid, coords, ell, sigma2=1.7, nugget=0.01, beta=0.8)
    X = rng.standard_normal((n, l))
    op = gsi.KernelCovMatrix(kind, coords, ell, sigma2=1.7, nugget=0.01, beta=0.8)
    assert op.shape == (n, n)
    assert op.T is op
    Y = op @ X
    assert relerr(Y, C @ X) < 1e-12
    assert relerr(X[:, :3].T @ op, X[:, :3].T @ C) < 1e-12


@pytest.mark.parametrize("kind", ["exponential", "gaussian", "powerlaw"])
@pytest.mark.parametrize("grid,spacing,l", [((40, 30), (1.0, 1.0), 60), ((14, 12, 10), (1.0, 0.5, 2.0), 210),
                                            ((1000,), (0.25,), 8), ((33, 31), (2.0, 3.0), 5), ((5, 1, 7), (1.0, 1.0, 1.0), 3)])
def test_grid_kernelcov_apply(gsi, kind, grid, spacing, l):
    """Structured-grid operator (lattice table look-up) vs the dense oracle and vs the
    arithmetic-generation operator on the same points."""
    rng = np.random.default_rng(len(grid) * 100 + l)
    coords = oracle.grid_coords(grid, spacing)
    d, n = coords.shape
    ell = [3.1, 2.7, 2.3][:d]
    kid = {"exponential": 0, "gaussian": 1, "powerlaw": 2}[kind]
    C = oracle.kernel_cov_dense(kid, coords, ell, sigma2=1.7, nugget=0.01, beta=0.8)
    X = rng.standard_normal((n, l))
    op = gsi.GridKernelCovMatrix(kind, grid, ell, spacing=spacing, sigma2=1.7, nugget=0.01, beta=0.8)
    assert op.shape == (n, n) and op.T is op
    Y = op @ X
    assert relerr(Y, C @ X) < 1e-12
    Yc = gsi.KernelCovMatrix(kind, coords, ell, sigma2=1.7, nugget=0.01, beta=0.8) @ X
    assert relerr(Y, Yc) < 1e-12


def test_kernelcov_unstructured_points(gsi):
    rng = np.random.default_rng(5)
    n = 1234
    coords = rng.uniform(0, 50, size=(3, n))
    ell = [9.0, 7.0, 5.0]
    C = oracle.kernel_cov_dense(0, coords, ell)
    X = rng.standard_normal((n, 27))
    assert relerr(gsi.KernelCovMatrix("exponential", coords, ell) @ X, C @ X) < 1e-12


@pytest.mark.parametrize("m,l", [(50, 7), (200, 60), (64, 64), (1000, 33), (5000, 210), (300, 256)])
def test_lu_L_matches_reference_rule(gsi, m, l):
    """`lu(Y).L` unpermuted (RandMatFact.jl:60-61): same pivots as LAPACK dgetrf."""
    rng = np.random.default_rng(m + l)
    Y = rng.standard_normal((m, l))
    L = gsi.lu_L(Y)
    Lref = oracle.lu_L_unpermuted(Y)
    assert L.shape == Lref.shape
    assert np.max(np.abs(L - Lref)) < 1e-11
    assert np.all(np.diag(L) == 1.0) and np.all(np.triu(L, 1) == 0.0)


def test_lu_tie_break_and_singular(gsi):
    Y = np.array([[1.0, 2.0], [-1.0, 0.5], [1.0, 3.0], [0.5, 1.0]])
    Lref, piv = gepp_L_unpermuted(Y)
    assert np.max(np.abs(gsi.lu_L(Y) - Lref)) < 1e-15
    Z = np.zeros((6, 3))
    Z[:, 0] = 1.0
    with pytest.raises(gsi.SingularException):
        gsi.lu_L(Z)


@pytest.mark.parametrize("m,l", [(10, 2), (100, 25), (64, 64), (3000, 210), (777, 110), (500, 256)])
def test_qr_thinQ(gsi, m, l):
    rng = np.random.default_rng(m * 3 + l)
    # graded columns so the conditioning is non trivial
    Y = rng.standard_normal((m, l)) * (10.0 ** (-8 * np.arange(l) / max(l - 1, 1)))[None, :]
    Q, R = gsi.qr_thinQ(Y, return_R=True)
    assert np.max(np.abs(Q.T @ Q - np.eye(l))) < 1e-13
    assert relerr(Q @ R, Y) < 1e-13
    assert np.all(np.tril(R, -1) == 0.0)
    Qo = oracle.randmatfact._qr_pivoted_thinQ(Y)
    # same range as the reference's pivoted QR (F2)
    assert np.linalg.norm(Qo - Q @ (Q.T @ Qo), 2) < 1e-6 * 1  # graded: range of tiny columns is ill-conditioned
    Y2 = rng.standard_normal((m, l))
    Q2 = gsi.qr_thinQ(Y2)
    Qo2 = oracle.randmatfact._qr_pivoted_thinQ(Y2)
    assert np.linalg.norm(Qo2 - Q2 @ (Q2.T @ Qo2), 2) < 1e-12


@pytest.mark.parametrize("l", [1, 2, 7, 60, 110, 210, 255, 256])
def test_svd_small(gsi, l):
    rng = np.random.default_rng(l)
    M = np.triu(rng.standard_normal((l, l))) * (10.0 ** (-6 * np.arange(l) / max(l - 1, 1)))[:, None]
    U, s = gsi.svd_small(M)
    sref = np.linalg.svd(M, compute_uv=False)
    # measured on B200 (tools/svd_accuracy.py -> profiles/r02/svd_drivers.json): <= 356 eps = 8e-14 with the block
    # ordering (this row-graded matrix takes 42 sweeps), 136 eps with the flat cyclic ordering
    assert np.max(np.abs(s - sref) / sref[0]) < 2e-13
    big = sref > 1e-9 * sref[0]
    assert np.max(np.abs(s[big] - sref[big]) / sref[big]) < 1e-9
    assert np.all(np.diff(s) <= 0)
    assert np.max(np.abs(U.T @ U - np.eye(l))) < 1e-12
    # U diag(s) V' = M  =>  U' M has rows of norm s
    assert np.max(np.abs(np.linalg.norm(U.T @ M, axis=1) - s) / sref[0]) < 1e-13


@pytest.mark.parametrize("l", [2, 3, 60, 129, 210, 256])
def test_svd_small_fused_matches_per_round(gsi, l):
    """The flat single-launch cluster driver of the Jacobi SVD ("svd.fused" = 2) performs the same
    rotations in the same order as the launch-per-round driver: bit-identical U and sigma.  The block
    driver (= 1, default: column blocks in shared memory, another cyclic ordering) agrees to rounding."""
    ctx = gsi.default_context()
    rng = np.random.default_rng(1000 + l)
    M = np.triu(rng.standard_normal((l, l))) * (10.0 ** (-6 * np.arange(l) / max(l - 1, 1)))[:, None]
    saved = ctx.get_option("svd.fused")
    try:
        ctx.set_option("svd.fused", 0)
        U0, s0 = gsi.svd_small(M)
        ctx.set_option("svd.fused", 2)
        U1, s1 = gsi.svd_small(M)
        ctx.set_option("svd.fused", 1)
        U2, s2 = gsi.svd_small(M)
    finally:
        ctx.set_option("svd.fused", saved)
    assert np.array_equal(s0, s1) and np.array_equal(U0, U1)
    assert np.max(np.abs(s2 - s0)) < 2e-13 * s0[0]
    big = s0 > 1e-9 * s0[0]
    assert np.max(np.abs(s2[big] - s0[big]) / s0[big]) < 1e-9
    assert np.max(np.abs(U2.T @ U2 - np.eye(l))) < 1e-12
    assert np.max(np.abs(np.linalg.norm(U2.T @ M, axis=1) - s2) / s0[0]) < 1e-13


@pytest.mark.parametrize("nobs", [7, 64, 200, 513])
def test_direct_solve_fused_matches_per_round(gsi, nobs):
    """Same for the pinv solve of pcgadirect (Jacobi with accumulated V, rectangular 2m x m):
    register-resident columns (2m <= 256), streamed columns, and the fall-back above 512 columns."""
    ctx = gsi.default_context()
    rng = np.random.default_rng(nobs)
    etas = [rng.standard_normal(nobs) for _ in range(5)]
    A = gsi.PCGALowRankMatrix(etas, rng.standard_normal(nobs), 1e-4 * (1 + rng.random(nobs)))
    b = np.concatenate([rng.standard_normal(nobs), [0.0]])
    saved = ctx.get_option("svd.fused")
    try:
        ctx.set_option("svd.fused", 0)
        x0 = A.pinv_solve(b)
        ctx.set_option("svd.fused", 2)
        x1 = A.pinv_solve(b)
        ctx.set_option("svd.fused", 1)
        x2, r2 = A.pinv_solve(b, return_rank=True)
        ctx.set_option("svd.fused", 0)
        _, r0 = A.pinv_solve(b, return_rank=True)
    finally:
        ctx.set_option("svd.fused", saved)
    assert np.array_equal(x0, x1)
    # the block ordering: same retained rank, same solution up to the conditioning of the saddle-point system
    assert r2 == r0
    assert np.linalg.norm(x2 - x0) <= 1e-6 * np.linalg.norm(x0)


def test_lowrankcov_algebra(gsi):
    """testrpcga.jl:46-58 on the device operator."""
    samples = [np.array([-.5, 0., .5]), np.array([1., -1., 0.]), np.array([-.5, 1., -.5])]
    lrcm = gsi.LowRankCovMatrix(samples)
    assert lrcm.shape == (3, 3) and lrcm.size(1) == 3
    with pytest.raises(ValueError):
        lrcm.size(3)
    fullcm = np.eye(3) @ lrcm
    assert np.allclose(fullcm, lrcm @ np.eye(3))
    assert np.allclose(sum(np.outer(x, x) for x in samples) / 2, fullcm)
    rng = np.random.default_rng(1)
    for _ in range(5):
        x = rng.standard_normal((3, 3))
        assert np.allclose(fullcm @ x, lrcm @ x)
        assert np.allclose(fullcm.T @ x, lrcm.T @ x)


def test_lowrankcov_vs_oracle(gsi):
    rng = np.random.default_rng(3)
    n, N, l = 2500, 100, 50
    fields = [rng.standard_normal(n) + 3.0 for _ in range(N)]
    X = rng.standard_normal((n, l))
    ref = oracle.LowRankCovMatrix(fields) @ X
    assert relerr(gsi.LowRankCovMatrix(fields) @ X, ref) < 1e-12


@pytest.mark.parametrize("table", [True, False])
def test_kernelcov_sweep_window(gsi, table):
    """The sweep window (gsi_ctx_set_option "kcov.window") only changes WHEN a CTA reads an X
    tile, never the order it accumulates them in: products are bit-identical with the window on
    and off, and match the dense oracle on sampled rows (1e-12).  Sized so that the persistent
    grid runs one full round plus a tail round (n > 64 * 148) with many 4-tile epochs."""
    ctx = gsi.default_context()
    grid, ell, l = (110, 109), [7.0, 5.0], 24
    coords = oracle.grid_coords(grid)
    n = coords.shape[1]
    X = np.random.default_rng(5).standard_normal((n, l))
    if table:
        op = gsi.GridKernelCovMatrix("exponential", grid, ell)
    else:
        op = gsi.KernelCovMatrix("exponential", coords, ell)
    saved = {k: ctx.get_option(k) for k in ("kcov.window", "kcov.epoch_shift", "kcov.sweep_div")}
    try:
        ctx.set_option("kcov.window", 0)
        ctx.set_option("kcov.sweep_div", -1)          # one tile between neighbouring sweep starts
        Y0 = op @ X
        ctx.set_option("kcov.epoch_shift", 2)
        for window in (1, 3):
            ctx.set_option("kcov.window", window)
            assert np.array_equal(op @ X, Y0)
    finally:
        for k, v in saved.items():
            ctx.set_option(k, v)
    rows = np.random.default_rng(6).choice(n, 200, replace=False)
    Cr = oracle.kernel_cov_dense(0, coords, ell, rows=rows)
    assert relerr(Y0[rows], Cr @ X) < 1e-12
    with pytest.raises(gsi.GsiError):
        ctx.set_option("kcov.sweep_groups", 3)
    with pytest.raises(gsi.GsiError):
        ctx.set_option("no.such.option", 1)


# ---------------------------------------------------------------- device FFTRF
@pytest.mark.parametrize("Ns", [(28, 37), (64, 32), (25, 50), (13, 17, 11), (32, 16, 8), (27, 31, 26)])
def test_fftrf_powerlaw_structuredgrid(gsi, Ns):
    """test/testfftrf.jl:6-15 (mean == k0, std == dk, size == Ns) and parity of the device sampler
    (direct DFT over the surviving outputs) with the oracle restatement of src/FFTRF.jl:83-100
    (FFT of the doubled grid) on identical phases."""
    from oracle.fftrf import powerlaw_structuredgrid as ref
    rng = np.random.default_rng(sum(Ns))
    k0, dk, beta = rng.standard_normal(), rng.random() + 0.1, -2 - rng.random()
    shape = gsi.FFTRF._doubled_shape(Ns)
    phi = rng.standard_normal(shape)
    k = gsi.FFTRF.powerlaw_structuredgrid(Ns, k0, dk, beta, phi=phi)
    assert list(k.shape) == list(Ns)
    assert abs(np.mean(k) - k0) < 1e-12 * max(1.0, abs(k0)) + 1e-12
    assert abs(np.std(k, ddof=1) - dk) < 1e-12
    kr = ref(list(Ns), k0, dk, beta, phi=phi)
    assert relerr(k, kr) < 1e-11


def test_getxis_device_sampler_matches_host_fields(gsi):
    """getxis(samplefield, numfields, numxis, p, q): the device sampler path (fields generated and kept on
    the device) returns the same fields and the same xis as handing the SAME fields to the host-list path
    (src/GeostatInversion.jl:29-38, 58-61)."""
    Ns, nf, K, p = (40, 36), 24, 8, 4
    n = Ns[0] * Ns[1]
    Omega = np.random.default_rng(1).standard_normal((n, K + p))
    sampler = gsi.PowerLawFieldSampler(Ns, 2.0, 3.14, -3.5, rng=11)
    xis, fields = gsi.getxis(sampler, nf, K, p, 3, Omega=Omega, want_fields=True)
    assert len(fields) == nf and len(xis) == K and fields[0].shape == (n,)
    it = iter(fields)
    xis_host = gsi.getxis(lambda: next(it), nf, K, p, 3, Omega=Omega)
    for a, b in zip(xis, xis_host):
        assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b) < 1e-10
    # the same phases through the oracle's FFT
    from oracle.fftrf import powerlaw_structuredgrid as ref
    rng = np.random.default_rng(11)
    phi = rng.standard_normal((nf,) + gsi.FFTRF._doubled_shape(Ns))
    assert relerr(fields[3], ref(list(Ns), 2.0, 3.14, -3.5, phi=phi[3]).ravel(order="F")) < 1e-11
