"""GPU parity tests of the PCGA / RGA path and of the remaining RandMatFact entry points,
through the C ABI, against the oracle (reference test/testrpcga.jl, test/testrmf.jl)."""
import numpy as np
import pytest

import oracle
from gpu_util import gsi, relerr  # noqa: F401

pytestmark = pytest.mark.gpu


def makeA(rng, n, m):
    return rng.standard_normal((n, m)) @ rng.standard_normal((m, n))


@pytest.mark.parametrize("n,m", [(10, 2), (10, 5), (100, 5), (100, 10), (100, 25)])
def test_rangefinder_adaptive(gsi, n, m):
    # test/testrmf.jl:13-15, and parity with the oracle on identical random vectors
    rng = np.random.default_rng(1000 * n + m)
    A = makeA(rng, n, m)
    Om0, oms = rng.standard_normal((n, 10)), rng.standard_normal((n, n))
    Q = gsi.rangefinder(A, Omega=Om0, omegas=oms)
    assert abs(Q.shape[1] - m) <= 1
    assert np.linalg.norm(A - Q @ Q.T @ A) < 1e-8
    Qo = oracle.rangefinder_adaptive(A, Om0, oms)
    assert Q.shape == Qo.shape
    assert np.linalg.norm(Qo - Q @ (Q.T @ Qo), 2) < 1e-8


def test_rangefinder_adaptive_wide(gsi):
    """the default call (maxvec = min(m, n) > 256 random vectors, COLMAJOR device buffers)
    on an operator with n > 256, parity with the oracle on identical vectors."""
    n, m = 600, 40
    rng = np.random.default_rng(77)
    A = makeA(rng, n, m)
    Om0, oms = rng.standard_normal((n, 10)), rng.standard_normal((n, n))
    Q = gsi.rangefinder(A, Omega=Om0, omegas=oms)
    assert abs(Q.shape[1] - m) <= 1
    assert np.linalg.norm(A - Q @ Q.T @ A) < 1e-8 * np.linalg.norm(A)
    Qo = oracle.rangefinder_adaptive(A, Om0, oms)
    assert Q.shape == Qo.shape
    assert np.linalg.norm(Qo - Q @ (Q.T @ Qo), 2) < 1e-8
    Q2 = gsi.rangefinder(A, rng=np.random.default_rng(1))          # vectors drawn inside, default maxvec
    assert abs(Q2.shape[1] - m) <= 1 and np.linalg.norm(A - Q2 @ Q2.T @ A) < 1e-8 * np.linalg.norm(A)


@pytest.mark.parametrize("n,m,block", [(100, 10, 4), (600, 40, 16), (3000, 120, 32), (1000, 300, 64)])
def test_rangefinder_adaptive_blocked(gsi, n, m, block):
    """the opt-in blocked adaptive range finder meets the reference's own acceptance
    test (test/testrmf.jl:13-15: ||A - QQ'A|| < 1e-8-ish, size ~ rank) with the basis size rounded
    up to the block, and spans the same subspace as its CPU statement on identical vectors."""
    rng = np.random.default_rng(n + m + block)
    A = makeA(rng, n, m)
    oms = rng.standard_normal((n, min(n, m + 4 * block)))
    Q = gsi.rangefinder(A, omegas=oms, block=block)
    assert m <= Q.shape[1] < m + 2 * block and Q.shape[1] % block == 0
    assert np.max(np.abs(Q.T @ Q - np.eye(Q.shape[1]))) < 1e-12
    assert np.linalg.norm(A - Q @ (Q.T @ A)) < 1e-8 * np.linalg.norm(A)
    Qo = oracle.rangefinder_adaptive_blocked(A, oms, block=block)
    assert Qo.shape == Q.shape
    # the first ceil(m / block) - 1 blocks are well conditioned: same subspace as the CPU statement
    lead = (m // block) * block if m % block else m - block
    if lead > 0:
        assert np.linalg.norm(Qo[:, :lead] - Q @ (Q.T @ Qo[:, :lead]), 2) < 1e-8
    with pytest.raises(gsi.GsiError):
        gsi.rangefinder(A, omegas=oms[:, :block], block=block)      # not enough vectors: NO_CONVERGENCE


def test_eig_nystrom_known_answer(gsi):
    # test/testrmf.jl:21-29: eigenvalues 2, 2 +- sqrt(2)
    rng = np.random.default_rng(7)
    A = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]])
    Q = gsi.rangefinder(A, rng=rng)
    U, Sigmavec = gsi.eig_nystrom(A, Q)
    expect = np.array([2 + np.sqrt(2), 2.0, 2 - np.sqrt(2)])
    assert np.linalg.norm(Sigmavec ** 2 - expect) < 1e-8
    assert np.max(np.abs(U.T @ U - np.eye(3))) < 1e-12


def test_eig_nystrom_vs_oracle_and_posdef_error(gsi):
    rng = np.random.default_rng(3)
    G = rng.standard_normal((300, 40))
    A = G @ G.T + 1e-3 * np.eye(300)
    Q = np.linalg.qr(rng.standard_normal((300, 25)))[0]
    U, S = gsi.eig_nystrom(A, Q)
    Uo, So = oracle.eig_nystrom(A, Q)
    assert np.max(np.abs(S - So) / So) < 1e-10
    assert np.linalg.norm(Uo - U @ (U.T @ Uo), 2) < 1e-8
    with pytest.raises(gsi.PosDefException):
        gsi.eig_nystrom(-A, Q)


def test_pcgalowrank_size_and_matvec(gsi):
    # testrpcga.jl:10-16 and :18-44
    rng = np.random.default_rng(5)
    A = gsi.PCGALowRankMatrix([rng.random(20) for _ in range(10)], rng.random(20), np.zeros(20))
    assert A.size() == (A.size(1), A.size(2)) == (21, 21)
    with pytest.raises(ValueError):
        A.size(3)
    numetas, numobs = 10, 20
    for noise in (1e16, 0.0):
        for etagen in ("zeros", "randn"):
            for hxgen in ("zeros", "randn"):
                gen = {"zeros": lambda k: np.zeros(k), "randn": lambda k: rng.standard_normal(k)}
                etas = [gen[etagen](numobs) for _ in range(numetas)]
                HX = gen[hxgen](numobs)
                R = noise * np.ones(numobs)
                lr = gsi.PCGALowRankMatrix(etas, HX, R)
                big = oracle.PCGALowRankMatrix(etas, HX, R).dense()
                for i in range(numobs + 1):
                    x = np.zeros(numobs + 1)
                    x[i] = 1.0
                    assert np.allclose(big @ x, lr @ x, rtol=np.sqrt(np.finfo(float).eps), atol=0)
    # dense R
    Rd = rng.standard_normal((numobs, numobs))
    Rd = Rd @ Rd.T
    etas = [rng.standard_normal(numobs) for _ in range(numetas)]
    HX = rng.standard_normal(numobs)
    x = rng.standard_normal(numobs + 1)
    assert relerr(gsi.PCGALowRankMatrix(etas, HX, Rd) @ x, oracle.PCGALowRankMatrix(etas, HX, Rd) @ x) < 1e-13


@pytest.mark.parametrize("nobs,K", [(20, 5), (200, 100), (500, 30)])
def test_device_lsqr_matches_oracle(gsi, nobs, K):
    """`IterativeSolvers.lsqr(bigA, b)` (src/lsqr.jl:54) on the device vs the oracle.

    With the package defaults (atol = btol = sqrt(eps)) LSQR stops while the iterate still
    carries ~1e-8..1e-5 relative error and a one-ulp rounding difference can move the stop
    by one iteration, so two correct implementations only agree to that level; with tight
    tolerances the same recurrence agrees to 1e-8 (measured 1e-9..1e-13)."""
    rng = np.random.default_rng(nobs + K)
    etas = [rng.standard_normal(nobs) for _ in range(K)]
    HX = rng.standard_normal(nobs)
    R = 1e-2 * np.ones(nobs)
    b = np.concatenate([rng.standard_normal(nobs), [0.0]])
    A, Ao = gsi.PCGALowRankMatrix(etas, HX, R), oracle.PCGALowRankMatrix(etas, HX, R)
    xe = np.linalg.lstsq(Ao.dense(), b, rcond=None)[0]
    x, info = A.lsqr(b, return_info=True)
    xo, infoo = oracle.lsqr(Ao, b, return_info=True)
    assert info["istop"] == infoo["istop"] and abs(info["itn"] - infoo["itn"]) <= 1, (info, infoo)
    assert relerr(x, xe) < 10 * max(relerr(xo, xe), 1e-9)        # same quality as the reference's stop
    tight = dict(atol=1e-14, btol=1e-14, conlim=1e16)
    x2, info2 = A.lsqr(b, return_info=True, **tight)
    xo2, infoo2 = oracle.lsqr(Ao, b, return_info=True, **tight)
    assert info2["istop"] == infoo2["istop"] and abs(info2["itn"] - infoo2["itn"]) <= 1
    assert relerr(x2, xo2) < 1e-8


@pytest.mark.parametrize("nobs,K,noise,dense_R", [(20, 5, 1e-1, False), (200, 100, 1e-2, False), (300, 30, 1e-4, True),
                                                  (64, 1, 1e-3, False), (513, 8, 1e-4, False)])
def test_device_direct_solve_matches_oracle(gsi, nobs, K, noise, dense_R):
    """`pinv([HQH + R, HX; HX', 0]) * b` (src/direct.jl:49-58) on the device vs the oracle's
    dgesdd pinv with Julia's cut-off.  Tolerance 10 * eps * cond (two SVD algorithms on one
    ill-conditioned matrix); the retained rank must be identical."""
    rng = np.random.default_rng(nobs * 7 + K)
    etas = [rng.standard_normal(nobs) for _ in range(K)]
    HX = rng.standard_normal(nobs)
    if dense_R:
        G = rng.standard_normal((nobs, nobs)) / np.sqrt(nobs)
        R = noise ** 2 * (np.eye(nobs) + 0.1 * (G + G.T))
    else:
        R = noise ** 2 * (1.0 + rng.random(nobs))
    b = np.concatenate([rng.standard_normal(nobs), [0.0]])
    big = oracle.PCGALowRankMatrix(etas, HX, R).dense()
    sv = np.linalg.svd(big, compute_uv=False)
    cut = np.finfo(float).eps * (nobs + 1) * sv[0]
    x, rank = gsi.PCGALowRankMatrix(etas, HX, R).pinv_solve(b, return_rank=True)
    assert rank == int(np.sum(sv > cut)) == nobs + 1
    cond = sv[0] / sv[-1]
    assert relerr(x, oracle.pinv(big) @ b) < 10 * np.finfo(float).eps * cond
    assert relerr(big @ x, b) < 20 * np.finfo(float).eps * cond


def test_device_direct_solve_rank_deficient(gsi):
    """pinv semantics: exactly singular system (HX = 0, R = 0, K < nobs) -> the null space is
    cut at eps * (nobs+1) * sigma_max and the minimum-norm solution is returned."""
    rng = np.random.default_rng(11)
    nobs, K = 40, 12
    etas = [rng.standard_normal(nobs) for _ in range(K)]
    HX, R = np.zeros(nobs), np.zeros(nobs)
    b = np.concatenate([rng.standard_normal(nobs), [0.0]])
    big = oracle.PCGALowRankMatrix(etas, HX, R).dense()
    x, rank = gsi.PCGALowRankMatrix(etas, HX, R).pinv_solve(b, return_rank=True)
    assert rank == K
    assert relerr(x, oracle.pinv(big) @ b) < 1e-10


from pcga_cases import setupsimpletest, TIGHT, SIMPLE_CASES  # noqa: E402
import pcga_cases as pc  # noqa: E402


@pytest.mark.parametrize("log2N,log2M,mu", SIMPLE_CASES)
def test_simpletestpcga(gsi, log2N, log2M, mu):
    """testrpcga.jl:125-131 end to end on the GPU path (2e-2 vs ground truth, the reference's own
    bar), plus parity with the oracle run on the SAME xis and the same host forward model.  Every
    bound is <= 10x the value measured on B200 (profiles/r02/pcga_parity_table.json):

    pcgalsqr, DEFAULT LSQR tolerances: the device LSQR stops at the same iteration with the same
      istop as the oracle and one iteration agrees to 1e-12 (measured <= 2.3e-14); with LSQR run
      to convergence 1e-8 (<= 1.5e-9); five-iteration default runs 2e-7 (<= 1.2e-8: each iteration
      re-draws ~1e-16/delta = 1e-8-relative rounding noise in every eta).
    pcgadirect: `pinv` (dgesdd) of a saddle-point matrix of condition 1e9..1e11 (R = 1e-8): one
      iteration agrees to eps * cond(bigA) (measured 0.1..0.4 eps cond = 7e-8..4e-6; dgesvd instead
      of dgesdd on the CPU moves the result as much), retained rank identical; full runs 6e-5
      (<= 5.6e-6)."""
    c = pc.simple_case(log2N, log2M, mu)
    forward, p0, X, R, yobs, truep, M = c["forward"], c["s0"], c["X"], c["R"], c["y"], c["truth"], c["K"]
    xis = gsi.getxis(c["Q"], M, c["p"], Omega=c["Omega"])
    delta = pc.DELTA
    from gsi_b200.pcga import pcgadirectiteration, pcgalsqriteration
    s1 = pcgadirectiteration(forward, p0, X, xis, R, yobs, delta, lambda s, o: None)
    s1o = oracle.pcgadirectiteration(forward, p0, X, xis, R, yobs, delta, lambda s, o: None)
    bigA = oracle.pcgadirect_system(forward, p0, X, xis, R, yobs, delta)[0]
    sv = np.linalg.svd(bigA, compute_uv=False)
    sv = sv[sv > np.finfo(float).eps * len(sv) * sv[0]]
    assert relerr(s1, s1o) < 4 * np.finfo(float).eps * sv[0] / sv[-1]
    popt = gsi.pcgadirect(forward, p0, X, xis, R, yobs)
    assert np.linalg.norm(popt - truep) / np.linalg.norm(truep) < 2e-2
    assert relerr(popt, oracle.pcgadirect(forward, p0, X, xis, R, yobs)) < 6e-5
    if c["lsqr_ok"]:
        assert pc.paramstorun_bit_identical(gsi, p0, X, xis)
        itg, ito, isg, iso, xrel = pc.lsqr_first_iteration_info(gsi, forward, p0, X, xis, R, yobs)
        assert (itg, isg) == (ito, iso) and xrel < 1e-12
        s1 = pcgalsqriteration(forward, p0, X, xis, R, yobs, delta)
        s1o = oracle.pcgalsqriteration(forward, p0, X, xis, R, yobs, delta)
        assert relerr(s1, s1o) < 1e-12                               # DEFAULT tolerances
        s1 = pcgalsqriteration(forward, p0, X, xis, R, yobs, delta, lsqr_kwargs=TIGHT)
        s1o = oracle.pcgalsqriteration(forward, p0, X, xis, R, yobs, delta, lsqr_kwargs=TIGHT)
        assert relerr(s1, s1o) < 1e-8
        popt = gsi.pcgalsqr(forward, p0, X, xis, R, yobs)
        assert np.linalg.norm(popt - truep) / np.linalg.norm(truep) < 2e-2
        assert relerr(popt, oracle.pcgalsqr(forward, p0, X, xis, R, yobs)) < 2e-7


def test_simpletestrga(gsi):
    # testrpcga.jl:133-138 (default pcgadirect) and the F5 case pcgafunc=pcgalsqr
    M, N, Nred, mu = 8, 1024, 512, 10.0
    rng = np.random.default_rng(N)
    forward, p0, X, Q, Omega, R, yobs, truep, pp = setupsimpletest(rng, M, N, mu)
    xis = gsi.getxis(Q, M, pp, Omega=Omega)
    S = rng.standard_normal((Nred, N)) * (1 / np.sqrt(N))
    calls = []
    popt = gsi.rga(forward, p0, X, xis, R, yobs, S, callback=lambda s, o: calls.append(len(o)))
    assert calls and all(c == Nred for c in calls)
    assert np.linalg.norm(popt - truep) / np.linalg.norm(truep) < 2e-2
    assert relerr(popt, oracle.rga(forward, p0, X, xis, R, yobs, S)) < 2e-4          # measured 1.1e-5
    popt2 = gsi.rga(forward, p0, X, xis, R, yobs, S, pcgafunc=gsi.pcgalsqr)
    assert np.linalg.norm(popt2 - truep) / np.linalg.norm(truep) < 2e-2
    pref2 = oracle.rga(forward, p0, X, xis, R, yobs, S, pcgafunc=oracle.pcgalsqr)
    assert relerr(popt2, pref2) < 1e-6      # measured 7.5e-8 (sketch-GEMM rounding amplified by 1/delta)


def test_pcga_more_than_253_xis(gsi):
    """the reference has no limit on the number of xis; the K+3 batch is a wide device iterate
    (> 256 columns) -- one iteration against the oracle on the same xis, black-box and declared linear model."""
    from gsi_b200.pcga import pcgalsqriteration, pcgadirectiteration, LinearForwardModel
    K, N, nobs = 300, 1500, 80
    rng = np.random.default_rng(300)
    Z = np.linalg.qr(rng.standard_normal((N, K)))[0] * (2.0 ** (-np.arange(K) / 40.0))[None, :]
    xis = [np.ascontiguousarray(Z[:, i]) for i in range(K)]
    H = rng.standard_normal((nobs, N)) / np.sqrt(N)
    truth = 1.5 + Z @ rng.standard_normal(K)
    R = 1e-2 * np.ones(nobs)                                    # well-conditioned saddle-point system: LSQR converges
    y = H @ truth + 1e-1 * rng.standard_normal(nobs)
    X, s0 = np.ones(N), np.full(N, 1.5)
    f = lambda s: H @ s                                         # noqa: E731
    assert pc.paramstorun_bit_identical(gsi, s0, X, xis)
    so = oracle.pcgalsqriteration(f, s0, X, xis, R, y, pc.DELTA, lsqr_kwargs=TIGHT)
    assert relerr(pcgalsqriteration(f, s0, X, xis, R, y, pc.DELTA, lsqr_kwargs=TIGHT), so) < 1e-7
    assert relerr(pcgalsqriteration(LinearForwardModel(H), s0, X, xis, R, y, pc.DELTA, lsqr_kwargs=TIGHT), so) < 1e-5
    sd = pcgadirectiteration(f, s0, X, xis, R, y, pc.DELTA, lambda s, o: None)
    sdo = oracle.pcgadirectiteration(f, s0, X, xis, R, y, pc.DELTA, lambda s, o: None)
    assert relerr(sd, sdo) < 1e-5


def test_rga_declared_linear_model_stays_on_device(gsi):
    """rga with a LinearForwardModel: H*P and S*(H*P) never visit the host; same estimate as the black-box
    path (whose batch is assembled in page-locked memory)."""
    from gsi_b200.pcga import LinearForwardModel
    M, N, Nred = 6, 900, 300
    rng = np.random.default_rng(77)
    H = rng.standard_normal((N, N)) / np.sqrt(N) + np.eye(N)
    Q = np.exp(-np.abs(np.subtract.outer(np.arange(N), np.arange(N))) / 40.0)
    Omega = rng.standard_normal((N, M + 10))
    xis = gsi.getxis(Q, M, 10, Omega=Omega)
    truep = 3.0 + np.stack(xis, axis=1) @ rng.standard_normal(M)
    R = 1e-6 * np.ones(N)
    yobs = H @ truep + 1e-3 * rng.standard_normal(N)
    X, p0 = np.ones(N), np.full(N, 3.0)
    S = rng.standard_normal((Nred, N)) / np.sqrt(N)
    pa = gsi.rga(LinearForwardModel(H), p0, X, xis, R, yobs, S, pcgafunc=gsi.pcgalsqr)
    pb = gsi.rga(lambda s: H @ s, p0, X, xis, R, yobs, S, pcgafunc=gsi.pcgalsqr)
    assert np.linalg.norm(pa - truep) / np.linalg.norm(truep) < 2e-2
    assert relerr(pa, pb) < 1e-5


def test_sketch_products(gsi):
    rng = np.random.default_rng(9)
    Nred, nobs = 300, 1500
    S = rng.standard_normal((Nred, nobs)) / np.sqrt(nobs)
    R = rng.random(nobs) + 0.1
    from gsi_b200.pcga import _Sketch
    sk = _Sketch(S, gsi.default_context())
    assert relerr(sk.cov(R), (S * R[None, :]) @ S.T) < 1e-13
    V = rng.standard_normal((nobs, 7))
    assert relerr(sk.apply(V), S @ V) < 1e-13
    y = rng.standard_normal(nobs)
    assert relerr(sk.apply(y), S @ y) < 1e-13
