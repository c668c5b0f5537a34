"""`RandMatFact` surface (reference src/RandMatFact.jl) over the B200 library.

Same names and argument meaning as the reference; the random matrix is drawn on the
host (seedable) and uploaded, exactly as the reference draws `randn(n, l)` on the host
(src/RandMatFact.jl:54) after `Random.seed!` (src/GeostatInversion.jl:24-27).
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import check, NORMALISER_LU_REF, LAYOUT_TALL, LAYOUT_COLMAJOR
from .core import DeviceMatrix, as_operator, _pd


def _rng(rng):
    if rng is None:
        return np.random.default_rng()
    if isinstance(rng, (int, np.integer)):
        return np.random.default_rng(int(rng))
    return rng


def _negative_iterations(q):
    # reference: error("parameter numiterations should be positive, but numiterations=$q")  :63
    raise ValueError(f"parameter numiterations should be positive, but numiterations={q}")


def rangefinder(A, l=None, numiterations=None, *, epsilon=1e-8, r=10, Omega=None, omegas=None, rng=None,
                normaliser=NORMALISER_LU_REF, maxvec=None, block=None):
    """rangefinder(A, l, numiterations)   -- fixed rank     (src/RandMatFact.jl:50-80)
    rangefinder(A; epsilon=1e-8, r=10) -- adaptive (4.2)   (src/RandMatFact.jl:15-48)
    rangefinder(A; epsilon, block=b)   -- OPT-IN blocked adaptive finder (not the parity mode: one
                                          GEMM pass over A per b vectors, see gsi_rangefinder_adaptive_blocked)"""
    op = as_operator(A)
    ctx = op.ctx
    m, n = op.shape
    rng = _rng(rng)
    if l is None and block is not None:
        return _rangefinder_adaptive_blocked(op, epsilon, int(block), omegas, rng, maxvec)
    if l is None:
        return _rangefinder_adaptive(op, epsilon, r, Omega, omegas, rng, maxvec)
    if numiterations is None:
        raise TypeError("rangefinder(A, l, numiterations): numiterations is required")
    if numiterations < 0:
        _negative_iterations(numiterations)
    if Omega is None:
        Omega = rng.standard_normal((n, l))
    Om, tmp = (Omega, False) if isinstance(Omega, DeviceMatrix) else (DeviceMatrix.from_host(ctx, Omega), True)
    Q = DeviceMatrix(ctx, op.mloc, l, LAYOUT_TALL)
    check(ctx._lib.gsi_rangefinder_fixed(op._h, Om._h, int(numiterations), int(normaliser), Q._h))
    if tmp:
        Om.free()
    out = Q.numpy()
    Q.free()
    return out


def _wide(ctx, a):
    """n x c host matrix -> device: TALL up to 256 columns, COLMAJOR beyond (any width)."""
    a = np.asarray(a, dtype=np.float64)
    return DeviceMatrix.from_host(ctx, a, LAYOUT_TALL if a.shape[1] <= 256 else LAYOUT_COLMAJOR)


def _rangefinder_adaptive(op, epsilon, r, Omega0, omegas, rng, maxvec):
    """The reference draws its vectors one by one for as long as it needs them (:36); here `maxvec` of
    them (default min(m, n), the most the algorithm can ever use) are drawn up front, in the
    reference's order of use, and uploaded as one matrix."""
    ctx = op.ctx
    m, n = op.shape
    if Omega0 is None:
        Omega0 = rng.standard_normal((n, r))
    if omegas is None:
        if maxvec is None:
            maxvec = min(m, n)
        omegas = rng.standard_normal((n, maxvec))
    maxvec = omegas.shape[1]
    O0 = DeviceMatrix.from_host(ctx, Omega0)
    Os = _wide(ctx, omegas)
    Q = DeviceMatrix(ctx, m, maxvec, Os.layout)
    j = C.c_int64()
    try:
        check(ctx._lib.gsi_rangefinder_adaptive(op._h, O0._h, Os._h, float(epsilon), int(r), Q._h, C.byref(j)))
        out = Q.numpy()[:, :j.value].copy()
    finally:
        O0.free(); Os.free(); Q.free()
    return out


def _rangefinder_adaptive_blocked(op, epsilon, block, omegas, rng, maxvec):
    ctx = op.ctx
    m, n = op.shape
    if omegas is None:
        if maxvec is None:
            maxvec = min(m, n)
        omegas = rng.standard_normal((n, maxvec))
    maxvec = omegas.shape[1]
    Os = _wide(ctx, omegas)
    Q = DeviceMatrix(ctx, m, maxvec, Os.layout)
    j = C.c_int64()
    try:
        check(ctx._lib.gsi_rangefinder_adaptive_blocked(op._h, Os._h, float(epsilon), int(block), Q._h, C.byref(j)))
        out = Q.numpy()[:, :j.value].copy()
    finally:
        Os.free(); Q.free()
    return out


def randsvd(A, K, p, q, *, Omega=None, rng=None, normaliser=NORMALISER_LU_REF, return_singular_values=False,
            device_out=False, full=False):
    """randsvd(A, K, p, q) -> Z = V*sqrt(S) (n x (K+p), last p columns exactly 0)
    (reference src/RandMatFact.jl:83-90).

    Omega (n x (K+p)) may be given (ndarray or DeviceMatrix); otherwise drawn from `rng`.
    On a multi-rank operator each rank gets its own row block of Z unless full=True.
    """
    op = as_operator(A)
    ctx = op.ctx
    m, n = op.shape
    K, p, q = int(K), int(p), int(q)
    if q < 0:
        _negative_iterations(q)
    l = K + p
    if Omega is None:
        Omega = _rng(rng).standard_normal((n, l))
    Om, tmp = (Omega, False) if isinstance(Omega, DeviceMatrix) else (DeviceMatrix.from_host(ctx, Omega), True)
    sharded_sym = ctx.world > 1 and op.symmetric and not full
    zrows = op.mloc if sharded_sym else n
    Z = DeviceMatrix(ctx, zrows, l, LAYOUT_TALL)
    S = np.empty(l, dtype=np.float64)
    try:
        check(ctx._lib.gsi_randsvd(op._h, Om._h, K, p, q, int(normaliser), Z._h, _pd(S)))
    finally:
        if tmp:
            Om.free()
    if device_out:
        return (Z, S) if return_singular_values else Z
    out = Z.numpy()
    Z.free()
    return (out, S) if return_singular_values else out


def eig_nystrom(A, Q):
    """eig_nystrom(A, Q) -> (U, Sigmavec)   (reference src/RandMatFact.jl:92-102)."""
    op = as_operator(A)
    ctx = op.ctx
    Qd = DeviceMatrix.from_host(ctx, Q)
    l = Qd.shape[1]
    U = DeviceMatrix(ctx, op.shape[0], l, LAYOUT_TALL)
    S = np.empty(l, dtype=np.float64)
    try:
        check(ctx._lib.gsi_eig_nystrom(op._h, Qd._h, U._h, _pd(S)))
        out = U.numpy()
    finally:
        Qd.free(); U.free()
    return out, S


def colnorms(Y):
    """src/RandMatFact.jl:7-13 (host helper; not on the device path)."""
    Y = np.asarray(Y)
    return np.sqrt(np.sum(Y * Y, axis=0))
