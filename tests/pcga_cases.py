"""Builders of the PCGA / RGA parity cases shared by tests/test_gpu_pcga*.py and
tools/pcga_parity_table.py: the cases of the reference's test/testrpcga.jl:104-138 and the
BASELINE configurations C2 / C4 at their stated sizes."""
import numpy as np
import scipy.linalg

import oracle
from oracle.fftrf import powerlaw_structuredgrid

DELTA = float(np.sqrt(np.finfo(float).eps))
TIGHT = dict(atol=1e-15, btol=1e-15, conlim=1e17)

# (log2N, log2M, mu) of test/testrpcga.jl:125-131 as run by tests/test_gpu_pcga.py
SIMPLE_CASES = [(4, 0, 0.0), (6, 2, 10.0), (8, 3, 0.0), (8, 5, 10.0), (8, 7, 0.0)]


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def setupsimpletest(rng, M, N, mu):
    """test/testrpcga.jl:104-123."""
    x = rng.standard_normal(N)
    Q0 = rng.standard_normal((M, N))
    Q = Q0.T @ Q0
    sqrtQ = np.real(scipy.linalg.sqrtm(Q))
    truep = sqrtQ @ rng.standard_normal(N) + mu
    forward = lambda p: p * x                          # noqa: E731  (elementwise model, :110-112)
    truey = forward(truep)
    pp = int(round(0.1 * M))
    Omega = rng.standard_normal((N, M + pp))
    X = np.full(N, mu)
    noiselevel = 0.0001
    R = noiselevel ** 2 * np.ones(N)
    yobs = truey + noiselevel * rng.standard_normal(N)
    p0 = np.full(N, mu)
    return forward, p0, X, Q, Omega, R, yobs, truep, pp


def simple_case(log2N, log2M, mu):
    N, M = 2 ** log2N, 2 ** log2M
    rng = np.random.default_rng(100 * log2N + log2M)
    forward, p0, X, Q, Omega, R, yobs, truep, pp = setupsimpletest(rng, M, N, mu)
    return dict(name=f"simple N=2^{log2N} M=2^{log2M} mu={mu:g}", forward=forward, s0=p0, X=X, Q=Q, Omega=Omega, R=R,
                y=yobs, truth=truep, K=M, p=pp, lsqr_ok=(M < N / 6))


def config2(full=True):
    """BASELINE configs[1]: 100x100 grid, exponential covariance ell = (12, 8), 200 synthetic linear
    observations, rank-100 prior (p = round(0.1 K) as in test/testrpcga.jl:114, q = 3 the getxis default)."""
    rng = np.random.default_rng(2)
    grid, nobs, K, p = ((100, 100), 200, 100, 10) if full else ((40, 40), 60, 40, 4)
    coords = oracle.grid_coords(grid)
    n = coords.shape[1]
    ell = [12.0, 8.0]
    Omega = rng.standard_normal((n, K + p))
    H = rng.standard_normal((nobs, n)) / np.sqrt(n)
    noise = 1e-4
    return dict(name=f"C2 {grid[0]}x{grid[1]} nobs={nobs} K={K}", grid=grid, coords=coords, ell=ell, n=n, nobs=nobs, K=K, p=p,
                q=3, Omega=Omega, H=H, mu=2.0, noise=noise, R=noise ** 2 * np.ones(nobs), X=np.full(n, 1.0),
                s0=np.full(n, 2.0), rng=rng)


def config2_truth(c, xis):
    Zk = np.stack(xis, axis=1)
    truth = c["mu"] + Zk @ c["rng"].standard_normal(c["K"])
    y = c["H"] @ truth + c["noise"] * c["rng"].standard_normal(c["nobs"])
    return truth, y


class GemmLowRankCov:
    """The reference's LowRankCovMatrix product A*B = S (S' B) / (N-1) (src/lowrank.jl:115-121) evaluated
    with two dgemm calls instead of N gemv + ger! pairs -- the loop form takes 40 s per randsvd at C4
    (profiles/r01/configs_c1_c2_c4.json); algebraically identical, checked against the loop form on a
    column subset by the C4 test."""
    __array_ufunc__ = None

    def __init__(self, samples):
        S = np.stack([np.asarray(s, dtype=np.float64) for s in samples], axis=1)
        means = np.zeros(S.shape[0])
        for i in range(S.shape[1]):                    # same order as src/lowrank.jl:19-24
            means += S[:, i]
        means = means / S.shape[1]
        self.S = np.asfortranarray(S - means[:, None])
        self.N = S.shape[1]

    @property
    def shape(self):
        return (self.S.shape[0], self.S.shape[0])

    @property
    def T(self):
        return self

    def __matmul__(self, B):
        return self.S @ (self.S.T @ np.asarray(B)) * (1.0 / (self.N - 1))

    def __rmatmul__(self, B):
        return (self @ np.asarray(B).T).T


def config4(full=True):
    """BASELINE configs[3]: 256x256 FFTRF power-law fields (k0 = 2, dk = 3.14, beta = -3.5, test/testrpcga.jl:87)
    -> LowRankCovMatrix prior; 10^5 elementwise observations h(s) = s[idx] .* x sketched to 500
    (S = randn(500, 1e5)/sqrt(1e5), test/testrpcga.jl:135); pcgalsqr on the sketched triple."""
    rng = np.random.default_rng(4)
    side, nf, nobs, Nred, K, p = (256, 200, 100000, 500, 30, 10) if full else (48, 40, 4000, 120, 12, 4)
    fields = [powerlaw_structuredgrid([side, side], 2.0, 3.14, -3.5, rng).ravel(order="F") for _ in range(nf)]
    n = side * side
    Omega = np.random.default_rng(5).standard_normal((n, K + p))
    idx = rng.integers(0, n, size=nobs)
    xmul = rng.standard_normal(nobs)
    forward = lambda s: s[idx] * xmul                  # noqa: E731
    S = rng.standard_normal((Nred, nobs)) / np.sqrt(nobs)
    noise = 1e-4
    mu = 2.0
    return dict(name=f"C4 {side}x{side} nf={nf} nobs={nobs}->{Nred} K={K}", fields=fields, n=n, nobs=nobs, Nred=Nred, K=K, p=p,
                q=3, Omega=Omega, forward=forward, S=S, noise=noise, R=noise ** 2 * np.ones(nobs), X=np.full(n, 1.0),
                s0=np.full(n, mu), mu=mu, rng=rng)


def config4_truth(c, xis):
    Zk = np.stack(xis, axis=1)
    truth = c["mu"] + Zk @ c["rng"].standard_normal(c["K"])
    y = c["forward"](truth) + c["noise"] * c["rng"].standard_normal(c["nobs"])
    return truth, y


def lsqr_first_iteration_info(gsi, forward, s, X, xis, R, y, delta=DELTA):
    """Runs the K+3 forward evaluations of one iteration on the host (identical on both sides), then the
    device LSQR and the oracle LSQR on the SAME saddle-point data with the package's default tolerances:
    returns (itn_gpu, itn_oracle, istop_gpu, istop_oracle, relerr of the LSQR solutions)."""
    K = len(xis)
    P = [s + delta * xi for xi in xis] + [s + delta * X, s + delta * s, s]
    res = [np.asarray(forward(pv), dtype=np.float64) for pv in P]
    hs = res[K + 2]
    etas = [(res[i] - hs) / delta for i in range(K)]
    HX = (res[K] - hs) / delta
    Hs = (res[K + 1] - hs) / delta
    b = np.concatenate([y - hs + Hs, np.zeros(1)])
    xg, ig = gsi.PCGALowRankMatrix(etas, HX, R).lsqr(b, return_info=True)
    xo, io = oracle.lsqr(oracle.PCGALowRankMatrix(etas, HX, R), b, return_info=True)
    return ig["itn"], io["itn"], ig["istop"], io["istop"], relerr(xg, xo)


def paramstorun_bit_identical(gsi, s, X, xis, delta=DELTA):
    """The device batch P = [s + delta*xi_i .., s + delta*X, s + delta*s, s] against the host expression."""
    from gsi_b200.pcga import _paramstorun, _xis_to_device
    ctx = gsi.default_context()
    Zk, K, tmp = _xis_to_device(ctx, xis)
    P = _paramstorun(ctx, Zk, K, s, X, delta)
    Ph = P.numpy()
    P.free()
    if tmp:
        Zk.free()
    ref = np.stack([s + delta * xi for xi in xis] + [s + delta * X, s + delta * s, s], axis=1)
    return bool(np.array_equal(Ph, ref))
