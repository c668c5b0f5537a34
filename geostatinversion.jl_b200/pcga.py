"""PCGA / RGA drivers and the prior-subspace builder behind the reference's keyword
surface (reference src/lsqr.jl:20-63, src/direct.jl:21-67, src/GeostatInversion.jl:20-103).

The user forward model is a host callable (black box), as in the reference where it
runs under `Distributed.pmap`.  Everything around it that is linear algebra -- the
`paramstorun` batch, the saddle-point LSQR solve with the matrix-free
PCGALowRankMatrix, the update s = X*beta + Z_K (E' xi), the rga sketch products, and,
for a *declared* linear forward model, the batched forward run H*P -- runs on the GPU.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import check, LAYOUT_TALL, LAYOUT_COLMAJOR, NORMALISER_LU_REF
from .core import (DeviceMatrix, LowRankCovMatrix, default_context, as_operator, _pd, _f64_colmajor, _Operator)
from . import randmatfact as _rmf

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


# ---------------------------------------------------------------- building blocks (tests / shim)
def lu_L(Y, ctx=None):
    """`lu(Y).L` with LAPACK row order kept (reference src/RandMatFact.jl:60-61)."""
    ctx = ctx or default_context()
    Yd = DeviceMatrix.from_host(ctx, Y)
    try:
        check(ctx._lib.gsi_lu_L(ctx._h, Yd._h))
        return Yd.numpy()
    finally:
        Yd.free()


def qr_thinQ(Y, ctx=None, return_R=False):
    """Thin orthonormal basis of range(Y) (replaces `Matrix(qr(Y, Val(true)).Q)`, :57-58)."""
    ctx = ctx or default_context()
    Yd = DeviceMatrix.from_host(ctx, Y)
    l = Yd.shape[1]
    R = np.zeros((l, l), order="F") if return_R else None
    try:
        check(ctx._lib.gsi_qr_thinQ(ctx._h, Yd._h, _pd(R) if return_R else None, l))
        Q = Yd.numpy()
    finally:
        Yd.free()
    return (Q, R) if return_R else Q


def svd_small(M, ctx=None):
    """SVD of a small square matrix on the device: returns (U, sigma)."""
    ctx = ctx or default_context()
    M = np.array(M, dtype=np.float64, order="F")
    l = M.shape[0]
    assert M.shape == (l, l)
    sig = np.empty(l)
    check(ctx._lib.gsi_svd_small(ctx._h, _pd(M), l, l, _pd(sig)))
    return M, sig


# ---------------------------------------------------------------- getxis
def getxis(Q, *args, ctx=None, rng=None, Omega=None, normaliser=NORMALISER_LU_REF, want_fields=False):
    """getxis(Q::Matrix, numxis, p, q=3, seed=nothing)                (GeostatInversion.jl:63-70)
    getxis(samplefield::Function, numfields, numxis, p, q=3, seed=nothing)   (:58-61, :29-38)

    `Q` may also be any gsi_b200 operator (e.g. KernelCovMatrix).  `seed` seeds the host
    generator that draws Omega (NumPy's, not Julia's stream).  Returns a list of numxis
    vectors (the first numxis columns of Z); with want_fields=True also the fields."""
    fields = None
    if callable(Q) and not isinstance(Q, _Operator):
        numfields, numxis, p = args[0], args[1], args[2]
        q = args[3] if len(args) > 3 else 3
        seed = args[4] if len(args) > 4 else None
        if hasattr(Q, "sample_device"):
            # device sampler (FFTRF.PowerLawFieldSampler): the fields never visit the host
            Sd = Q.sample_device(numfields)
            if want_fields:                                                        # before the mean is removed in place
                Sh = Sd.numpy()
                fields = [np.ascontiguousarray(Sh[:, i]) for i in range(Sh.shape[1])]
            A = LowRankCovMatrix.from_device(Sd)
        else:
            fields = [np.asarray(Q(), dtype=np.float64) for _ in range(numfields)]     # rpmap(i->samplefield())
            A = LowRankCovMatrix(fields, ctx=ctx)
    else:
        numxis, p = args[0], args[1]
        q = args[2] if len(args) > 2 else 3
        seed = args[3] if len(args) > 3 else None
        A = as_operator(Q, ctx)
    if rng is None and seed is not None:
        rng = np.random.default_rng(seed)
    Z = _rmf.randsvd(A, numxis, p, q, Omega=Omega, rng=rng, normaliser=normaliser)
    xis = [np.ascontiguousarray(Z[:, i]) for i in range(numxis)]
    if want_fields:
        return xis, fields
    return xis


# ---------------------------------------------------------------- PCGALowRankMatrix
def _split_R(R, nobs):
    """-> (Rdiag or None, Rdense or None)."""
    if np.isscalar(R):
        return np.full(nobs, float(R)), None
    if hasattr(R, "todense") or hasattr(R, "toarray"):          # scipy.sparse
        import scipy.sparse as sp
        Rs = sp.csr_matrix(R)
        d = Rs.diagonal()
        if (Rs - sp.diags(d)).nnz == 0:
            return np.ascontiguousarray(d, dtype=np.float64), None
        return None, np.asfortranarray(Rs.toarray(), dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 1:
        return np.ascontiguousarray(R), None
    return None, np.asfortranarray(R)


class PCGALowRankMatrix:
    """[HQH' + R, HX; HX', 0] with HQH' = sum_i eta_i eta_i' applied matrix-free on the
    device (reference src/lowrank.jl:32-36, mul! :83-97, size :62-73, adjoint :38-44)."""
    __array_ufunc__ = None

    def __init__(self, etas, HX, R, ctx=None):
        self.ctx = ctx or default_context()
        if isinstance(etas, (list, tuple)):
            E = np.stack([np.asarray(e, dtype=np.float64) for e in etas], axis=1)
        else:
            E = np.asarray(etas, dtype=np.float64)
        self.E = np.asfortranarray(E)
        self.HX = np.ascontiguousarray(HX, dtype=np.float64)
        self.nobs, self.K = self.E.shape
        self.Rdiag, self.Rdense = _split_R(R, self.nobs)

    @property
    def shape(self):
        return (self.nobs + 1, self.nobs + 1)

    def size(self, i=None):
        if i is None:
            return self.shape
        if i in (1, 2):
            return self.nobs + 1
        raise ValueError(f"there is no {i}-th dimension in a PCGALowRankMatrix")

    @property
    def T(self):
        return self

    def _rargs(self):
        return (_pd(self.Rdiag) if self.Rdiag is not None else None,
                _pd(self.Rdense) if self.Rdense is not None else None, self.nobs)

    def __matmul__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.empty(self.nobs + 1)
        rd, rD, ldr = self._rargs()
        check(self.ctx._lib.gsi_pcga_lowrank_matvec(self.ctx._h, self.nobs, self.K, _pd(self.E), self.nobs,
                                                    _pd(self.HX), rd, rD, ldr, _pd(x), _pd(v)))
        return v

    def lsqr(self, b, atol=0.0, btol=0.0, conlim=0.0, maxiter=0, return_info=False):
        """`IterativeSolvers.lsqr(bigA, b)` (src/lsqr.jl:54) run entirely on the device."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.nobs + 1)
        itn, istop = C.c_int64(), C.c_int32()
        rd, rD, ldr = self._rargs()
        check(self.ctx._lib.gsi_pcga_lsqr_solve(self.ctx._h, self.nobs, self.K, _pd(self.E), self.nobs, _pd(self.HX),
                                                rd, rD, ldr, _pd(b), atol, btol, conlim, maxiter, _pd(x),
                                                C.byref(itn), C.byref(istop)))
        if return_info:
            return x, dict(itn=itn.value, istop=istop.value)
        return x

    def pinv_solve(self, b, return_rank=False):
        """`pinv(bigA) * b` for the DENSE saddle-point matrix [HQH + R, HX; HX', 0] that
        pcgadirect assembles (src/direct.jl:49-58), on the device (gsi_pcga_direct_solve)."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.nobs + 1)
        rank = C.c_int64()
        rd, rD, ldr = self._rargs()
        check(self.ctx._lib.gsi_pcga_direct_solve(self.ctx._h, self.nobs, self.K, _pd(self.E), self.nobs,
                                                  _pd(self.HX), rd, rD, ldr, _pd(b), _pd(x), C.byref(rank)))
        return (x, rank.value) if return_rank else x


# ---------------------------------------------------------------- forward models
class LinearForwardModel:
    """A *declared* linear forward model h(s) = H s.  Called with a host vector it
    behaves like any black-box model; `apply_batch` runs all K+3 runs of a PCGA
    iteration as ONE tensor-core GEMM H * P on the device."""

    def __init__(self, H, ctx=None):
        self.ctx = ctx or default_context()
        self.H = np.asfortranarray(H, dtype=np.float64)
        from .core import DenseMatrix
        self.op = DenseMatrix(self.H, self.ctx)

    def __call__(self, s):
        return self.op.apply(np.asarray(s, dtype=np.float64))

    def apply_batch_device(self, P):
        """P: TALL DeviceMatrix n x c  ->  TALL DeviceMatrix nobs x c (caller frees)."""
        return self.op.apply(P)

    def apply_batch(self, P):
        """P: TALL DeviceMatrix n x c  ->  host array nobs x c."""
        Y = self.apply_batch_device(P)
        out = Y.numpy()
        Y.free()
        return out


MAX_XIS = 1021     # the paramstorun batch [xis.., X, s, s] is one device iterate of at most 1024 columns


def _xis_to_device(ctx, xis):
    if isinstance(xis, DeviceMatrix):
        return xis, xis.shape[1], False
    if len(xis) > MAX_XIS:
        raise ValueError(f"pcgalsqr / pcgadirect / rga take at most {MAX_XIS} xis on the device path "
                         f"(the batch of K+3 parameter vectors is one device iterate of <= 1024 columns); got {len(xis)}")
    Zk = np.stack([np.asarray(x, dtype=np.float64) for x in xis], axis=1)
    return DeviceMatrix.from_host(ctx, Zk), Zk.shape[1], True


def _paramstorun(ctx, Zk, K, s, X, delta):
    P = DeviceMatrix(ctx, Zk.shape[0], K + 3, LAYOUT_TALL)
    s = np.ascontiguousarray(s, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64)
    check(ctx._lib.gsi_pcga_paramstorun(ctx._h, Zk._h, K, _pd(s), _pd(X), float(delta), P._h))
    return P


def _run_forward(forwardmodel, P, pmap):
    if hasattr(forwardmodel, "apply_batch"):
        res = forwardmodel.apply_batch(P)
        return [np.ascontiguousarray(res[:, i]) for i in range(res.shape[1])]
    Ph = P.numpy()
    return [np.asarray(r, dtype=np.float64) for r in pmap(forwardmodel, [np.ascontiguousarray(Ph[:, i])
                                                                        for i in range(Ph.shape[1])])]


def _finish_iteration(ctx, Zk, K, X, R, y, results, delta, solver):
    hs = results[K + 2]
    E = np.empty((len(hs), K), order="F")
    for i in range(K):
        E[:, i] = (results[i] - hs) / delta                     # etas          lsqr.jl:46-49
    HX = (results[K] - hs) / delta                              # :50
    Hs = (results[K + 1] - hs) / delta                          # :51
    b = np.concatenate([y - hs + Hs, np.zeros(1)])              # :52
    x = solver(E, HX, R, b)
    s = np.empty(Zk.shape[0])
    X = np.ascontiguousarray(X, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    check(ctx._lib.gsi_pcga_update(ctx._h, Zk._h, K, _pd(X), _pd(E), E.shape[0], E.shape[0], _pd(x), _pd(s)))
    return s


def pcgalsqriteration(forwardmodel, s, X, xis, R, y, delta, callback=None, ctx=None, pmap=map, _dev=None,
                      lsqr_kwargs=None):
    """One PCGA/LSQR iteration (reference src/lsqr.jl:35-63).  lsqr_kwargs (atol, btol,
    conlim, maxiter) default to the IterativeSolvers defaults the reference relies on."""
    ctx = ctx or default_context()
    Zk, K, tmp = _dev if _dev is not None else _xis_to_device(ctx, xis)
    try:
        P = _paramstorun(ctx, Zk, K, s, X, delta)               # :37-43
        results = _run_forward(forwardmodel, P, pmap)           # :44  (pmap)
        P.free()
        if callback is not None:
            callback(s, results[K + 2])

        def solver(E, HX, R_, b):
            return PCGALowRankMatrix(E, HX, R_, ctx).lsqr(b, **(lsqr_kwargs or {}))    # :53-54

        return _finish_iteration(ctx, Zk, K, X, R, np.asarray(y, dtype=np.float64), results, delta, solver)
    finally:
        if tmp and _dev is None:
            Zk.free()


def _outer(iteration, s0, maxiters, xtol):
    converged = False
    s = np.asarray(s0, dtype=np.float64)
    itercount = 0
    while not converged and itercount < maxiters:              # lsqr.jl:24-31 / direct.jl:26-33
        olds = s
        s = iteration(s)
        if np.linalg.norm(s - olds) < xtol:
            converged = True
        itercount += 1
    return s


def pcgalsqr(forwardmodel, s0, X, xis, R, y, maxiters=5, delta=SQRT_EPS, xtol=1e-6, callback=None, ctx=None,
             pmap=map, lsqr_kwargs=None):
    """pcgalsqr(forwardmodel, s0, X, xis, R, y; maxiters=5, delta=sqrt(eps), xtol=1e-6)
    (reference src/lsqr.jl:20-33).  Also accepts `callback` (the reference's rga always
    forwards one)."""
    ctx = ctx or default_context()
    dev = _xis_to_device(ctx, xis)
    try:
        return _outer(lambda s: pcgalsqriteration(forwardmodel, s, X, xis, R, y, delta, callback, ctx, pmap, dev,
                                                   lsqr_kwargs),
                      s0, maxiters, xtol)
    finally:
        if dev[2]:
            dev[0].free()


def pcgadirectiteration(forwardmodel, s, X, xis, R, y, delta, callback, ctx=None, pmap=map, _dev=None):
    """One direct PCGA iteration (reference src/direct.jl:37-67).  The (nobs+1)^2
    saddle-point system is assembled and solved on the device with the reference's `pinv`
    semantics (tensor-core E E', one-sided Jacobi SVD, Julia's rtol cut-off)."""
    ctx = ctx or default_context()
    Zk, K, tmp = _dev if _dev is not None else _xis_to_device(ctx, xis)
    try:
        P = _paramstorun(ctx, Zk, K, s, X, delta)
        results = _run_forward(forwardmodel, P, pmap)
        P.free()
        callback(s, results[K + 2])                             # direct.jl:47

        def solver(E, HX, R_, b):
            # HQH = sum_i eta_i eta_i' (ger! loop, :49-53), bigA (:57), pinv(bigA) * b (:58)
            return PCGALowRankMatrix(E, HX, R_, ctx).pinv_solve(b)

        return _finish_iteration(ctx, Zk, K, X, R, np.asarray(y, dtype=np.float64), results, delta, solver)
    finally:
        if tmp and _dev is None:
            Zk.free()


def pcgadirect(forwardmodel, s0, X, xis, R, y, maxiters=5, delta=SQRT_EPS, xtol=1e-6,
               callback=lambda s, obs_cal: None, ctx=None, pmap=map):
    """pcgadirect(...; maxiters, delta, xtol, callback) (reference src/direct.jl:21-35)."""
    ctx = ctx or default_context()
    dev = _xis_to_device(ctx, xis)
    try:
        return _outer(lambda s: pcgadirectiteration(forwardmodel, s, X, xis, R, y, delta, callback, ctx, pmap, dev),
                      s0, maxiters, xtol)
    finally:
        if dev[2]:
            dev[0].free()


pcga = pcgadirect                                               # GeostatInversion.jl:105


class _Sketch:
    """Device-resident sketch matrix S (Nred x nobs) for rga."""

    def __init__(self, S, ctx):
        self.ctx = ctx
        self.S = np.asfortranarray(S, dtype=np.float64)
        self.buf = DeviceMatrix.from_host(ctx, self.S, LAYOUT_COLMAJOR)
        self._batch = None

    def apply_device(self, Vd):
        """S * V for a TALL device V (nobs x c) -> host array Nred x c."""
        out = DeviceMatrix(self.ctx, self.S.shape[0], Vd.shape[1], LAYOUT_TALL)
        try:
            check(self.ctx._lib.gsi_sketch_apply(self.ctx._h, self.buf._h, Vd._h, out._h))
            return out.numpy()
        finally:
            out.free()

    def apply(self, V):
        """S * V for host V (nobs x c or vector)."""
        V2 = _f64_colmajor(V)
        Vd = DeviceMatrix.from_host(self.ctx, V2)
        try:
            res = self.apply_device(Vd)
        finally:
            Vd.free()
        return res[:, 0] if np.ndim(V) == 1 else res

    def batch_buffer(self, nobs, c):
        """Page-locked column-major staging array for the nobs x c batch of forward runs of one
        iteration (reused across iterations: the upload is then a plain DMA, no bounce copy)."""
        if self._batch is None or self._batch.shape != (nobs, c):
            self._batch = self.ctx.pinned_empty((nobs, c))
        return self._batch

    def cov(self, R):
        """S * R * S'."""
        nobs = self.S.shape[1]
        rd, rD = _split_R(R, nobs)
        Nred = self.S.shape[0]
        if rd is None:
            # dense R: S * (R * S') as two sketch products
            return self.apply(np.asfortranarray(rD @ self.S.T))
        out = np.empty((Nred, Nred), order="F")
        check(self.ctx._lib.gsi_sketch_cov(self.ctx._h, self.buf._h, _pd(rd), _pd(out), Nred))
        return out


class _SketchedModel:
    def __init__(self, forwardmodel, sk):
        self.f, self.sk = forwardmodel, sk

    def __call__(self, x):
        return self.sk.apply(self.f(x))                         # x -> S * forwardmodel(x)

    def apply_batch(self, P):
        if hasattr(self.f, "apply_batch_device"):
            # declared linear model: H*P stays on the device, S*(H*P) follows without a host round trip
            Vd = self.f.apply_batch_device(P)
            try:
                return self.sk.apply_device(Vd)
            finally:
                Vd.free()
        if hasattr(self.f, "apply_batch"):
            return self.sk.apply(self.f.apply_batch(P))
        Ph = P.numpy()
        V = None
        for i in range(Ph.shape[1]):                            # assembled column-major in page-locked memory
            v = np.asarray(self.f(np.ascontiguousarray(Ph[:, i])), dtype=np.float64)
            if V is None:
                V = self.sk.batch_buffer(v.shape[0], Ph.shape[1])
            V[:, i] = v
        return self.sk.apply(V)                                 # all K+3 sketches as one GEMM


def rga(forwardmodel, s0, X, xis, R, y, S, maxiters=5, delta=SQRT_EPS, xtol=1e-6, pcgafunc=None,
        callback=lambda s, obs_cal: None, ctx=None):
    """rga(forwardmodel, s0, X, xis, R, y, S; maxiters, delta, xtol, pcgafunc=pcgadirect, callback)
    (reference src/GeostatInversion.jl:101-103): `x->S*h(x)`, `S*R*S'`, `S*y`."""
    ctx = ctx or default_context()
    if pcgafunc is None:
        pcgafunc = pcgadirect
    sk = _Sketch(S, ctx)
    return pcgafunc(_SketchedModel(forwardmodel, sk), s0, X, xis, sk.cov(R), sk.apply(np.asarray(y, dtype=np.float64)),
                    maxiters=maxiters, delta=delta, xtol=xtol, callback=callback, ctx=ctx)
