"""The drivers of the LU / QR panel factorisations must agree: the single-launch panel kernels
(panel rows resident in shared memory; option value 1 = a single thread-block cluster for short
iterates / cooperative grid otherwise, 2 = always the cooperative grid) against the
launch-per-column drivers of the first round (0), and all against the oracle's LAPACK calls
(reference src/RandMatFact.jl:60-61,75-76 -> dgetrf / dgeqp3)."""
import numpy as np
import pytest

import oracle
from gpu_util import gsi, relerr  # noqa: F401

pytestmark = pytest.mark.gpu


def _with_option(ctx, name, value, fn):
    saved = ctx.get_option(name)
    try:
        ctx.set_option(name, value)
        return fn()
    finally:
        ctx.set_option(name, saved)


# (n, l): single CTA; several CTAs; odd panel remainders; the widest iterate; more rows per CTA than
# shared memory holds (overflow rows worked on in place: n > 148 * 1432)
LU_SHAPES = [(64, 8), (40, 33), (300, 17), (1000, 60), (5000, 110), (20000, 210), (777, 256), (2049, 16),
             (230000, 40), (3000, 300), (1500, 520),      # the last two: iterates wider than 256 columns
             (10000, 110), (25088, 48), (26900, 20), (27100, 20),   # one cluster of 16 CTAs; rows beyond its shared memory; just past it
             (300000, 20)]                                          # cooperative grid with rows beyond the SMs' shared memory


@pytest.mark.parametrize("n,l", LU_SHAPES)
def test_lu_panel_matches_per_column_and_oracle(gsi, n, l):
    from gsi_b200.pcga import lu_L
    ctx = gsi.default_context()
    Y = np.random.default_rng(n + l).standard_normal((n, l))
    L0 = _with_option(ctx, "lu.panel", 0, lambda: lu_L(Y))
    L1 = _with_option(ctx, "lu.panel", 1, lambda: lu_L(Y))
    L2 = _with_option(ctx, "lu.panel", 2, lambda: lu_L(Y))
    assert np.array_equal(L0, L1)                       # same pivots, same arithmetic
    assert np.array_equal(L0, L2)
    assert relerr(L1, oracle.lu_L_unpermuted(Y)) < 1e-10


def test_lu_panel_ties_and_zero_pivot(gsi):
    """LAPACK tie rule (first row of maximal |value|) and the SingularException mapping."""
    from gsi_b200.pcga import lu_L
    ctx = gsi.default_context()
    Y = np.ones((200, 6))
    Y[:, 1] = np.arange(200) % 7
    Y[:, 2] = -(np.arange(200) % 5)
    Y[:, 3:] = np.random.default_rng(0).integers(-3, 4, size=(200, 3))

    def run():
        try:
            return lu_L(Y), None
        except gsi.SingularException as e:
            return None, e
    L1, e1 = _with_option(ctx, "lu.panel", 1, run)
    L2, e2 = _with_option(ctx, "lu.panel", 2, run)
    L0, e0 = _with_option(ctx, "lu.panel", 0, run)
    assert (e0 is None) == (e1 is None) == (e2 is None)
    if e0 is None:
        assert np.array_equal(L0, L1) and np.array_equal(L0, L2)
    else:
        assert str(e0) == str(e1) == str(e2)
    # integer-valued columns with many exact ties, non-singular
    rng = np.random.default_rng(1)
    T = rng.integers(-2, 3, size=(3000, 24)).astype(np.float64) + 8.0 * np.eye(3000, 24)
    L1 = _with_option(ctx, "lu.panel", 1, lambda: lu_L(T))
    assert np.array_equal(L1, _with_option(ctx, "lu.panel", 0, lambda: lu_L(T)))
    assert np.array_equal(L1, _with_option(ctx, "lu.panel", 2, lambda: lu_L(T)))
    assert relerr(L1, oracle.lu_L_unpermuted(T)) < 1e-12


def test_lu_exact_zero_pivot_raises(gsi):
    from gsi_b200.pcga import lu_L
    Y = np.random.default_rng(2).standard_normal((500, 20))
    Y[:, 7] = 0.0
    for panel in (1, 2, 0):
        with pytest.raises(gsi.SingularException):
            _with_option(gsi.default_context(), "lu.panel", panel, lambda: lu_L(Y))
    with pytest.raises(oracle.randmatfact.SingularException):
        oracle.lu_L_unpermuted(Y)


def test_lu_wide_matrix(gsi):
    """n < l: L is n x n unit lower triangular (Julia's F.L is m x min(m, n)); the columns beyond n are zero."""
    from gsi_b200.pcga import lu_L
    Y = np.random.default_rng(5).standard_normal((24, 50))
    L = lu_L(Y)
    assert relerr(L[:, :24], oracle.lu_L_unpermuted(Y)) < 1e-12
    assert np.all(L[:, 24:] == 0.0)


def test_lu_nan_propagates(gsi):
    """A NaN in a pivot column is chosen as the pivot and propagates (it does not leave stale pivots)."""
    from gsi_b200.pcga import lu_L
    Y = np.random.default_rng(3).standard_normal((1000, 12))
    Y[417, 3] = np.nan
    masks = []
    for panel in (1, 2, 0):
        L = _with_option(gsi.default_context(), "lu.panel", panel, lambda: lu_L(Y))
        assert np.isnan(L).any()
        masks.append(np.isnan(L))
    assert np.array_equal(masks[0], masks[1]) and np.array_equal(masks[0], masks[2])


QR_SHAPES = [(64, 8), (1000, 60), (20000, 210), (300, 256), (5000, 33), (2049, 16), (230000, 24), (3000, 300), (1500, 520),
             (10000, 110), (25088, 48), (26900, 20), (27100, 20), (300000, 20)]


@pytest.mark.parametrize("n,l", QR_SHAPES)
def test_qr_panel(gsi, n, l):
    """Orthonormal Q spanning range(Y), QR = Y, R equal to the per-column driver's to rounding, and
    range-equivalent to the oracle's dgeqp3 Q."""
    from gsi_b200.pcga import qr_thinQ
    ctx = gsi.default_context()
    rng = np.random.default_rng(n * 3 + l)
    Y = rng.standard_normal((n, l)) * (10.0 ** (-4 * np.arange(l) / max(l - 1, 1)))[None, :]
    Q0, R0 = _with_option(ctx, "qr.panel", 0, lambda: qr_thinQ(Y, return_R=True))
    Q1, R1 = _with_option(ctx, "qr.panel", 1, lambda: qr_thinQ(Y, return_R=True))
    Q1b = _with_option(ctx, "qr.panel", 1, lambda: qr_thinQ(Y))
    assert np.array_equal(Q1, Q1b)                                  # deterministic
    Q2, R2 = _with_option(ctx, "qr.panel", 2, lambda: qr_thinQ(Y, return_R=True))     # cooperative-grid transport
    assert np.max(np.abs(Q2.T @ Q2 - np.eye(l))) < 1e-12
    assert relerr(Q2 @ R2, Y) < 1e-13
    assert relerr(R2, R0) < 1e-11 and relerr(Q2, Q0) < 1e-9
    assert np.max(np.abs(Q1.T @ Q1 - np.eye(l))) < 1e-12
    assert relerr(Q1 @ R1, Y) < 1e-13
    assert relerr(R1, R0) < 1e-11 and relerr(Q1, Q0) < 1e-9
    if n <= 20000:
        Qo = oracle.randmatfact._qr_pivoted_thinQ(Y)
        assert np.linalg.norm(Qo - Q1 @ (Q1.T @ Qo), 2) < 1e-10


def test_randsvd_same_result_with_either_driver(gsi):
    grid, ell, K, p, q = (40, 30), [6.0, 4.0], 40, 5, 2
    coords = oracle.grid_coords(grid)
    Omega = np.random.default_rng(3).standard_normal((coords.shape[1], K + p))
    op = gsi.KernelCovMatrix("exponential", coords, ell)
    ctx = gsi.default_context()
    Zp = gsi.randsvd(op, K, p, q, Omega=Omega)
    Zc = _with_option(ctx, "lu.panel", 0, lambda: _with_option(ctx, "qr.panel", 0,
                                                               lambda: gsi.randsvd(op, K, p, q, Omega=Omega)))
    Zl = _with_option(ctx, "qr.panel", 0, lambda: gsi.randsvd(op, K, p, q, Omega=Omega))
    assert np.array_equal(_with_option(ctx, "lu.panel", 0, lambda: gsi.randsvd(op, K, p, q, Omega=Omega)), Zp)
    assert np.array_equal(Zl, Zc)
    c = oracle.compare_Z(Zp, Zc, K)
    assert c["sv_rel"] < 1e-12 and c["sine"] < 1e-10
    Zref = oracle.randsvd(oracle.kernel_cov_dense(0, coords, ell), Omega, K, p, q)
    c = oracle.compare_Z(Zp, Zref, K)
    assert c["tail_zero"] and c["sv_rel"] < 1e-10 and c["sine"] < 1e-8


def test_gaussian_kernel_far_apart_points(gsi):
    """exp(-y) of the arithmetic generation path for y >> 2^31 ln2/64 (points thousands of
    length scales apart) must be 0, not a wrapped exponent."""
    n = 4096
    x = np.zeros((1, n))
    x[0] = np.arange(n) * 50.0                  # extent / ell = 2e5 length scales
    ell = [1.0]
    X = np.random.default_rng(0).standard_normal((n, 8))
    for kind, kid in (("gaussian", 1), ("exponential", 0)):
        op = gsi.KernelCovMatrix(kind, x * (1e3 if kind == "gaussian" else 1e6), ell)
        Y = op @ X
        assert np.all(np.isfinite(Y))
        assert relerr(Y, X) < 1e-14             # C == I to working precision


@pytest.mark.parametrize("K,p,q,dense", [(290, 10, 2, False), (290, 10, 1, True), (560, 40, 1, False)])
def test_randsvd_wider_than_256_columns(gsi, K, p, q, dense):
    """randsvd(A, K, p, q) has no width limit in the reference (src/RandMatFact.jl:83): iterates with
    K + p > 256 columns run the products in 256-column chunks and the factorisations on the wide buffer."""
    grid, ell = (64, 40), [9.0, 6.0]
    coords = oracle.grid_coords(grid)
    n = coords.shape[1]
    C = oracle.kernel_cov_dense(0, coords, ell)
    Omega = np.random.default_rng(K).standard_normal((n, K + p))
    op = gsi.DenseMatrix(C) if dense else gsi.KernelCovMatrix("exponential", coords, ell)
    Z, S = gsi.randsvd(op, K, p, q, Omega=Omega, return_singular_values=True)
    assert Z.shape == (n, K + p)
    c = oracle.compare_Z(Z, oracle.randsvd(C, Omega, K, p, q), K)
    assert c["tail_zero"] and c["sv_rel"] < 1e-10 and c["sine"] < 1e-8, c
    X = np.random.default_rng(1).standard_normal((n, 300))
    assert relerr(op @ X, C @ X) < 1e-12                    # the operator itself on a wide host matrix
