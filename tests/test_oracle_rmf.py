"""Pins the oracle against the reference's own RandMatFact tests
(reference test/testrmf.jl) -- CPU only."""
import numpy as np
import pytest

import oracle
from oracle.gepp_ref import gepp_L_unpermuted


def makeA(rng, n, m):
    # test/testrmf.jl:5-9
    return rng.standard_normal((n, m)) @ rng.standard_normal((m, n))


@pytest.mark.parametrize("n,m", [(10, 2), (10, 5), (100, 5), (100, 10), (100, 25)])  # testrmf.jl:32-36
def test_rangefinder(n, m):
    rng = np.random.default_rng(1000 * n + m)
    A = makeA(rng, n, m)
    # adaptive: testrmf.jl:13-15
    Q = oracle.rangefinder_adaptive(A, rng.standard_normal((n, 10)), rng.standard_normal((n, n)))
    assert abs(Q.shape[1] - m) <= 1
    assert np.linalg.norm(A - Q @ Q.T @ A) < 1e-8
    # fixed rank, q = 2: testrmf.jl:16-18
    Q = oracle.rangefinder_fixed(A, rng.standard_normal((n, m)), 2)
    assert abs(Q.shape[1] - m) <= 1
    assert np.linalg.norm(A - Q @ Q.T @ A) < 1e-8


def test_eig_nystrom_known_answer():
    # testrmf.jl:21-29 -- the only closed-form KAT: eigenvalues 2, 2 +- sqrt(2)
    rng = np.random.default_rng(7)
    A = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]])
    Q = oracle.rangefinder_adaptive(A, rng.standard_normal((3, 10)), rng.standard_normal((3, 3)))
    U, Sigmavec = oracle.eig_nystrom(A, Q)
    Lam = Sigmavec * Sigmavec
    expect = np.array([2 + np.sqrt(2), 2.0, 2 - np.sqrt(2)])
    assert np.linalg.norm(Lam - expect) < 1e-8
    assert np.linalg.norm(np.sort(np.linalg.eigvalsh(A))[::-1] - Lam) < 1e-8


def test_negative_iterations_errors():
    rng = np.random.default_rng(0)
    A = makeA(rng, 10, 2)
    with pytest.raises(ValueError, match="numiterations should be positive"):
        oracle.rangefinder_fixed(A, rng.standard_normal((10, 2)), -1)   # RandMatFact.jl:63


def test_q0_is_pivoted_qr_range():
    rng = np.random.default_rng(3)
    A = makeA(rng, 60, 8)
    Q = oracle.rangefinder_fixed(A, rng.standard_normal((60, 8)), 0)
    assert np.allclose(Q.T @ Q, np.eye(8), atol=1e-13)
    assert np.linalg.norm(A - Q @ Q.T @ A) < 1e-8


@pytest.mark.parametrize("m,n", [(50, 7), (200, 60), (64, 64), (1000, 33)])
def test_lu_L_matches_independent_gepp(m, n):
    rng = np.random.default_rng(m + n)
    Y = rng.standard_normal((m, n))
    L = oracle.lu_L_unpermuted(Y)
    Lref, piv = gepp_L_unpermuted(Y)
    assert L.shape == (m, min(m, n))
    assert np.max(np.abs(L - Lref)) < 1e-12
    assert np.all(np.diag(L) == 1.0) and np.all(np.triu(L, 1) == 0.0)
    assert np.max(np.abs(L)) <= 1.0 + 1e-15           # partial pivoting bound


def test_lu_L_is_NOT_unpermuted_range():
    """range(L) = P range(Y) != range(Y) in general."""
    rng = np.random.default_rng(11)
    Y = rng.standard_normal((40, 5))
    L = oracle.lu_L_unpermuted(Y)
    Qy, _ = np.linalg.qr(Y)
    Ql, _ = np.linalg.qr(L)
    assert np.linalg.norm(Ql - Qy @ (Qy.T @ Ql), 2) > 1e-2


def test_lu_tie_break_first_max():
    Y = np.array([[1.0, 2.0], [-1.0, 0.5], [1.0, 3.0], [0.5, 1.0]])
    L, piv = gepp_L_unpermuted(Y)
    assert piv[0] == 0                                  # first of the tied |1.0| rows
    assert np.max(np.abs(oracle.lu_L_unpermuted(Y) - L)) < 1e-15


def test_lu_singular_raises():
    Y = np.zeros((6, 3))
    Y[:, 0] = 1.0
    with pytest.raises(ArithmeticError):
        oracle.lu_L_unpermuted(Y)


def test_randsvd_structure_and_accuracy():
    rng = np.random.default_rng(2017)
    n, r, K, p = 300, 20, 20, 5
    A = makeA(rng, n, r)
    Z = oracle.randsvd(A, rng.standard_normal((n, K + p)), K, p, 2)
    assert Z.shape == (n, K + p)
    assert np.all(Z[:, K:] == 0.0)                      # RandMatFact.jl:87
    s = oracle.singvals_from_Z(Z, K)
    sref = np.linalg.svd(A, compute_uv=False)[:K]
    assert np.max(np.abs(s - sref) / sref) < 1e-10
