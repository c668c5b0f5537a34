"""Oracle restatement of the PCGA drivers (reference: src/lsqr.jl,
src/direct.jl, src/GeostatInversion.jl).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
`Distributed.pmap(forwardmodel, paramstorun)` is a plain serial map here (the
reference's own tests run it with zero workers, test/testrpcga.jl:149).
"""
import numpy as np
from .lowrank import PCGALowRankMatrix, LowRankCovMatrix
from .lsqr import lsqr as _lsqr
from . import randmatfact as _rmf

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


def _paramstorun(s, X, xis, delta):
    """src/lsqr.jl:37-43 == src/direct.jl:39-45."""
    P = [s + delta * xi for xi in xis]
    P.append(s + delta * X)
    P.append(s + delta * s)
    P.append(s)
    return P


def _Radd(HQH, R):
    if np.isscalar(R):
        return HQH + R * np.eye(HQH.shape[0])
    if hasattr(R, "ndim") and R.ndim == 1:
        return HQH + np.diag(R)
    if hasattr(R, "todense"):
        return HQH + np.asarray(R.todense())
    return HQH + np.asarray(R)


def pcgalsqriteration(forwardmodel, s, X, xis, R, y, delta, callback=None, lsqr_kwargs=None):
    """src/lsqr.jl:35-63."""
    K = len(xis)
    results = [np.asarray(forwardmodel(pv), dtype=np.float64) for pv in _paramstorun(s, X, xis, delta)]  # :44
    hs = results[K + 2]
    if callback is not None:
        callback(s, hs)
    etas = [(results[i] - hs) / delta for i in range(K)]   # :46-49
    HX = (results[K] - hs) / delta                         # :50
    Hs = (results[K + 1] - hs) / delta                     # :51
    b = np.concatenate([y - hs + Hs, np.zeros(1)])         # :52
    bigA = PCGALowRankMatrix(etas, HX, R)                  # :53
    x = _lsqr(bigA, b, **(lsqr_kwargs or {}))              # :54
    beta_bar = x[-1]
    xi_bar = x[:-1]
    snew = X * beta_bar                                    # :57
    for i in range(K):                                     # :58-61
        etai = (results[i] - hs) / delta
        snew = snew + xis[i] * np.dot(etai, xi_bar)
    return snew


def pcgalsqr(forwardmodel, s0, X, xis, R, y, maxiters=5, delta=SQRT_EPS, xtol=1e-6,
             callback=None, lsqr_kwargs=None):
    """src/lsqr.jl:20-33 (+ optional `callback`)."""
    converged = False
    s = np.asarray(s0, dtype=np.float64)
    itercount = 0
    while not converged and itercount < maxiters:
        olds = s
        s = pcgalsqriteration(forwardmodel, s, X, xis, R, y, delta, callback, lsqr_kwargs)
        if np.linalg.norm(s - olds) < xtol:
            converged = True
        itercount += 1
    return s


def pcgadirect_system(forwardmodel, s, X, xis, R, y, delta, callback=lambda s, obs: None):
    """src/direct.jl:39-57: the K+3 forward runs and the dense saddle-point system.
    Returns (bigA, b, E) with E = [eta_1 .. eta_K] as columns."""
    K = len(xis)
    results = [np.asarray(forwardmodel(pv), dtype=np.float64) for pv in _paramstorun(s, X, xis, delta)]
    callback(s, results[K + 2])                            # :47
    hs = results[K + 2]
    nobs = len(y)
    HQH = np.zeros((nobs, nobs))                           # :49
    E = np.empty((nobs, K))
    for i in range(K):                                     # :50-53
        etai = (results[i] - hs) / delta
        E[:, i] = etai
        HQH += np.outer(etai, etai)
    HX = (results[K] - hs) / delta
    Hs = (results[K + 1] - hs) / delta
    b = np.concatenate([y - hs + Hs, np.zeros(1)])         # :56
    bigA = np.block([[_Radd(HQH, R), HX[:, None]], [HX[None, :], np.zeros((1, 1))]])  # :57
    return bigA, b, E


def pinv(A):
    """Julia `pinv(A)` with its defaults: atol = 0, rtol = eps * min(size(A)) (dgesdd SVD)."""
    return np.linalg.pinv(A, rcond=np.finfo(np.float64).eps * min(A.shape))


def pcgadirectiteration(forwardmodel, s, X, xis, R, y, delta, callback):
    """src/direct.jl:37-67."""
    K = len(xis)
    bigA, b, E = pcgadirect_system(forwardmodel, s, X, xis, R, y, delta, callback)
    x = pinv(bigA) @ b                                     # :58
    beta_bar = x[-1]
    xi_bar = x[:-1]
    snew = X * beta_bar
    for i in range(K):
        snew = snew + xis[i] * np.dot(E[:, i], xi_bar)
    return snew


def pcgadirect(forwardmodel, s0, X, xis, R, y, maxiters=5, delta=SQRT_EPS, xtol=1e-6,
               callback=lambda s, obs: None):
    """src/direct.jl:21-35."""
    converged = False
    s = np.asarray(s0, dtype=np.float64)
    itercount = 0
    while not converged and itercount < maxiters:
        olds = s
        s = pcgadirectiteration(forwardmodel, s, X, xis, R, y, delta, callback)
        if np.linalg.norm(s - olds) < xtol:
            converged = True
        itercount += 1
    return s


def _RSt(R, S):
    """S * R * S' for scalar / diagonal-vector / dense / scipy-sparse R."""
    if np.isscalar(R):
        return R * (S @ S.T)
    if hasattr(R, "ndim") and R.ndim == 1:
        return (S * R[None, :]) @ S.T
    return S @ (R @ S.T)


def rga(forwardmodel, s0, X, xis, R, y, S, maxiters=5, delta=SQRT_EPS, xtol=1e-6,
        pcgafunc=pcgadirect, callback=lambda s, obs: None):
    """src/GeostatInversion.jl:101-103."""
    return pcgafunc(lambda x: S @ forwardmodel(x), s0, X, xis, _RSt(R, S), S @ y,
                    maxiters=maxiters, delta=delta, xtol=xtol, callback=callback)


def getxis(Q, Omega, numxis, p, q=3):
    """src/GeostatInversion.jl:63-70 / :29-38 with Omega replacing the seeded
    `randn` (randsvdwithseed, :20-27).  Q: dense matrix or LowRankCovMatrix."""
    Z = _rmf.randsvd(Q, Omega, numxis, p, q)
    return [Z[:, i].copy() for i in range(numxis)]
