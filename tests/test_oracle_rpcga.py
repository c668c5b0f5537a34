"""Pins the oracle against the reference's PCGA / operator tests
(reference test/testrpcga.jl) -- CPU only."""
import numpy as np
import pytest
import scipy.linalg

import oracle
from oracle.fftrf import powerlaw_structuredgrid


def test_pcgalowranksize():
    # testrpcga.jl:10-16
    rng = np.random.default_rng(0)
    A = oracle.PCGALowRankMatrix([rng.random(20) for _ in range(10)], rng.random(10), 0.0)
    assert A.shape == (A.size(1), A.size(2)) == (21, 21)
    with pytest.raises(ValueError):
        A.size(3)


@pytest.mark.parametrize("noise", [1e16, 0.0])
@pytest.mark.parametrize("etagen", ["zeros", "randn"])
@pytest.mark.parametrize("hxgen", ["zeros", "randn"])
def test_simplepcgalowrank(noise, etagen, hxgen):
    # testrpcga.jl:18-44
    rng = np.random.default_rng(5)
    numetas, numobs = 10, 20
    gen = {"zeros": lambda k: np.zeros(k), "randn": lambda k: rng.standard_normal(k)}
    etas = [gen[etagen](numobs) for _ in range(numetas)]
    HX = gen[hxgen](numobs)
    R = noise * np.ones(numobs)
    lr = oracle.PCGALowRankMatrix(etas, HX, R)
    big = lr.dense()
    for i in range(numobs + 1):
        x = np.zeros(numobs + 1)
        x[i] = 1.0
        assert np.allclose(big @ x, lr @ x, rtol=np.sqrt(np.finfo(float).eps), atol=0)


def test_simplelowrankcov():
    # testrpcga.jl:46-58
    samples = [np.array([-.5, 0., .5]), np.array([1., -1., 0.]), np.array([-.5, 1., -.5])]
    lrcm = oracle.LowRankCovMatrix(samples)
    fullcm = np.eye(3) @ lrcm
    assert np.allclose(fullcm, lrcm @ np.eye(3))
    assert np.allclose(sum(np.outer(x, x) for x in samples) / (len(samples) - 1), fullcm)
    rng = np.random.default_rng(1)
    for _ in range(100):
        x = rng.standard_normal((3, 3))
        assert np.allclose(fullcm @ x, lrcm @ x)
        assert np.allclose(fullcm.T @ x, lrcm.T @ x)


def test_lowrankcovconsistency():
    # testrpcga.jl:60-81 (N reduced 10000 -> 2000 to keep the CPU suite fast)
    rng = np.random.default_rng(2)
    N, M = 2000, 100
    sqrtcov = rng.standard_normal((M, M))
    cov = sqrtcov @ sqrtcov.T
    samples = [sqrtcov @ rng.standard_normal(M) for _ in range(N)]
    lrcm = oracle.LowRankCovMatrix(samples)
    full = lrcm @ np.eye(M)
    assert np.linalg.norm(full - cov, 2) <= M ** 2 / np.sqrt(N) + 10
    for _ in range(20):
        x = rng.standard_normal(M)
        assert np.allclose(lrcm @ x, full @ x)


def test_lowrankcovgetxis():
    # testrpcga.jl:83-102: getxis on the operator vs on its dense materialisation,
    # K=30 p=20 q=3, same Omega: equal up to sign at 1e-6.
    rng = np.random.default_rng(0)
    numfields, numxis, p = 100, 30, 20
    fields = [powerlaw_structuredgrid([25, 25], 2.0, 3.14, -3.5, rng).ravel(order="F")
              for _ in range(numfields)]
    lrcm = oracle.LowRankCovMatrix(fields)
    full = np.eye(625) @ lrcm
    Omega = np.random.default_rng(0).standard_normal((625, numxis + p))
    lrxis = oracle.getxis(lrcm, Omega, numxis, p, 3)
    fullxis = oracle.getxis(full, Omega, numxis, p, 3)
    for a, b in zip(fullxis, lrxis):
        assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 1e-6


def setupsimpletest(rng, M, N, mu):
    # testrpcga.jl:104-123
    x = rng.standard_normal(N)
    Q0 = rng.standard_normal((M, N))
    Q = Q0.T @ Q0
    sqrtQ = np.real(scipy.linalg.sqrtm(Q))
    truep = sqrtQ @ rng.standard_normal(N) + mu
    forward = lambda p: p * x
    truey = forward(truep)
    pp = int(round(0.1 * M))
    Omega = rng.standard_normal((N, M + pp))
    xis = oracle.getxis(Q, Omega, M, pp)
    X = np.full(N, mu)
    noiselevel = 0.0001
    R = noiselevel ** 2 * np.ones(N)
    yobs = truey + noiselevel * rng.standard_normal(N)
    p0 = np.full(N, mu)
    return forward, p0, X, xis, R, yobs, truep


@pytest.mark.parametrize("log2N,log2M,mu", [(l2n, l2m, mu) for l2n in (2, 4, 6, 8)
                                            for l2m in range(0, l2n, 2) for mu in (0.0, 10.0)])
def test_simpletestpcga(log2N, log2M, mu):
    # testrpcga.jl:125-131, 160-171 (subsampled grid of (M, N))
    N, M = 2 ** log2N, 2 ** log2M
    rng = np.random.default_rng(100 * log2N + log2M)
    forward, p0, X, xis, R, yobs, truep = setupsimpletest(rng, M, N, mu)
    popt = oracle.pcgadirect(forward, p0, X, xis, R, yobs)
    assert np.linalg.norm(popt - truep) / np.linalg.norm(truep) < 2e-2
    if M < N / 6:
        popt = oracle.pcgalsqr(forward, p0, X, xis, R, yobs)
        assert np.linalg.norm(popt - truep) / np.linalg.norm(truep) < 2e-2


@pytest.mark.parametrize("N,mu", [(2 ** 10, 0.0), (2 ** 10, 10.0)])
def test_simpletestrga(N, mu):
    # testrpcga.jl:133-138, 156-159
    M, Nred = 8, 512
    rng = np.random.default_rng(N + int(mu))
    forward, p0, X, xis, R, yobs, truep = setupsimpletest(rng, M, N, mu)
    S = rng.standard_normal((Nred, N)) * (1 / np.sqrt(N))
    popt = oracle.rga(forward, p0, X, xis, R, yobs, S)
    assert np.linalg.norm(popt - truep) / np.linalg.norm(truep) < 2e-2
    # F5: rga with pcgalsqr == pcgalsqr on the sketched triple
    popt2 = oracle.rga(forward, p0, X, xis, R, yobs, S, pcgafunc=oracle.pcgalsqr)
    assert np.linalg.norm(popt2 - truep) / np.linalg.norm(truep) < 2e-2


def test_lsqr_matches_scipy():
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(9)
    A = rng.standard_normal((40, 25))
    b = rng.standard_normal(40)
    x = oracle.lsqr(A, b)
    xs = spla.lsqr(A, b, atol=1.49e-8, btol=1.49e-8, conlim=1 / 1.49e-8, iter_lim=40)[0]
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-6
    assert np.linalg.norm(x - np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-6


def test_pinv_cutoff_is_julias():
    """oracle.pinv = Julia `pinv(A)`: SVD with singular values <= eps*min(size)*sigma_max
    dropped (atol = 0).  Checked on a matrix with a prescribed spectrum straddling the cut-off."""
    rng = np.random.default_rng(4)
    m = 30
    U, _ = np.linalg.qr(rng.standard_normal((m, m)))
    V, _ = np.linalg.qr(rng.standard_normal((m, m)))
    eps = np.finfo(float).eps
    sig = np.ones(m)
    sig[-3:] = [1e-6, 1e-3 * eps * m, 0.0]                   # kept, dropped, dropped
    A = (U * sig) @ V.T
    P = oracle.pinv(A)
    keep = sig > eps * m * sig.max()
    assert keep.sum() == m - 2
    Pexp = (V[:, keep] / sig[keep]) @ U[:, keep].T
    assert np.linalg.norm(P - Pexp) / np.linalg.norm(Pexp) < 1e-8
    assert np.linalg.matrix_rank(P, tol=1e-3) == m - 2


def test_pcgadirect_system_is_what_the_iteration_solves():
    """pcgadirect_system (src/direct.jl:39-57) + pinv (:58) + update (:59-65) == pcgadirectiteration,
    and bigA equals the dense form of the matrix-free PCGALowRankMatrix the LSQR path uses."""
    rng = np.random.default_rng(6)
    N, M = 32, 4
    x = rng.standard_normal(N)
    forward = lambda p: p * x
    xis = [rng.standard_normal(N) for _ in range(M)]
    X, s0 = np.ones(N), np.full(N, 2.0)
    R = 1e-8 * np.ones(N)
    y = forward(rng.standard_normal(N) + 2.0)
    delta = float(np.sqrt(np.finfo(float).eps))
    bigA, b, E = oracle.pcgadirect_system(forward, s0, X, xis, R, y, delta)
    assert bigA.shape == (N + 1, N + 1) and np.array_equal(bigA, bigA.T)
    HX = bigA[:N, N]
    assert np.allclose(bigA, oracle.PCGALowRankMatrix(list(E.T), HX, R).dense(), rtol=1e-14, atol=0)
    xsol = oracle.pinv(bigA) @ b
    s1 = X * xsol[-1]
    for i in range(M):
        s1 = s1 + xis[i] * np.dot(E[:, i], xsol[:-1])
    assert np.array_equal(s1, oracle.pcgadirectiteration(forward, s0, X, xis, R, y, delta, lambda s, o: None))
