"""Oracle restatement of `RandMatFact` (reference: src/RandMatFact.jl).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function follows the
cited reference lines statement by statement and calls the LAPACK routine that
Julia's LinearAlgebra dispatches to.  `A` may be a NumPy matrix or any object
with `.shape`, `@` (A @ X), and `.T` whose result supports `@` -- the same
duck-typed surface the reference requires (src/RandMatFact.jl:52-55,67,70,85).
"""
import numpy as np
from scipy.linalg import lapack, qr as _scipy_qr, svd as _scipy_svd


class SingularException(ArithmeticError):
    """Julia's LinearAlgebra.SingularException (lu(...; check=true))."""

    def __init__(self, info):
        super().__init__(f"SingularException({info})")
        self.info = info


class PosDefException(ArithmeticError):
    """Julia's LinearAlgebra.PosDefException (cholesky(...; check=true))."""

    def __init__(self, info):
        super().__init__(f"PosDefException({info})")
        self.info = info


def colnorms(Y):
    """src/RandMatFact.jl:7-13 -- 2-norm of every column."""
    Y = np.asarray(Y)
    return np.array([np.linalg.norm(Y[:, i]) for i in range(Y.shape[1])])


def lu_L_unpermuted(Y):
    """`F = lu(Y); Q = F.L` of src/RandMatFact.jl:60-61,68-69,72-73.

    Julia's `lu` is LAPACK dgetrf (partial pivoting, check=true).  `F.L` is the
    unit-lower-trapezoidal factor of Y[p, :] -- the reference uses it WITHOUT
    un-permuting.  Returned as an m x min(m, n) matrix.
    """
    Y = np.array(Y, dtype=np.float64, order="F")
    lu, piv, info = lapack.dgetrf(Y)
    if info < 0:
        raise ValueError(f"dgetrf illegal argument {-info}")
    if info > 0:
        raise SingularException(info)
    m, n = Y.shape
    k = min(m, n)
    L = np.tril(lu[:, :k], -1)
    L[np.arange(k), np.arange(k)] = 1.0
    return L


def _qr_pivoted_thinQ(Y):
    """`F = qr(Y, Val(true)); Matrix(F.Q)` (src/RandMatFact.jl:57-58,75-76):
    LAPACK dgeqp3 followed by dorgqr (thin Q)."""
    Q, _, _ = _scipy_qr(np.asarray(Y, dtype=np.float64), mode="economic", pivoting=True)
    return Q


def _adjoint(A):
    return A.T


def rangefinder_fixed(A, Omega, numiterations):
    """src/RandMatFact.jl:50-80 with `Omega = randn(n, l)` (line 54) supplied."""
    l = Omega.shape[1]
    n = A.shape[1]
    assert Omega.shape[0] == n
    Y = A @ Omega                                       # :55
    if numiterations == 0:
        return _qr_pivoted_thinQ(Y)                     # :56-58
    elif numiterations > 0:
        Q = lu_L_unpermuted(Y)                          # :59-61
    else:
        raise ValueError(                               # :62-64 (Julia `error(...)`)
            f"parameter numiterations should be positive, but numiterations={numiterations}")
    for i in range(1, numiterations + 1):               # :66
        Q = _adjoint(A) @ Q                             # :67
        Q = lu_L_unpermuted(Q)                          # :68-69
        Q = A @ Q                                       # :70
        if i < numiterations:
            Q = lu_L_unpermuted(Q)                      # :71-73
        else:
            Q = _qr_pivoted_thinQ(Q)                    # :74-76
    return Q


def randsvd(A, Omega, K, p, q):
    """src/RandMatFact.jl:83-90.  Returns Z (n x (K+p)); last p columns are 0."""
    assert Omega.shape[1] == K + p
    Q = rangefinder_fixed(A, Omega, q)                  # :84
    B = Q.T @ A                                         # :85  (operators: __rmatmul__)
    _, S, Vt = _scipy_svd(np.asarray(B), full_matrices=False, lapack_driver="gesdd")  # :86
    Sh = np.sqrt(np.concatenate([S[:K], np.zeros(p)]))  # :87
    Z = Vt.T * Sh[None, :]                              # :88
    return Z


def rangefinder_adaptive(A, Omega0, omegas, epsilon=1e-8, r=10):
    """src/RandMatFact.jl:15-48 (Halko et al. Alg 4.2).

    Omega0 (n x r) replaces `randn(n, r)` (line 20); column t of `omegas`
    replaces the t-th `randn!(omega)` (line 36).
    """
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    assert Omega0.shape == (n, r)
    Yfull = np.zeros((n, r + min(n, m)))                # :18 (n rows: square A only)
    Yfull[:, :r] = A @ Omega0                           # :19-20
    j = 0
    Qfull = np.zeros((m, min(n, m)))                    # :23
    thresh = epsilon / np.sqrt(200 / np.pi)
    while np.max(colnorms(Yfull[:, j:j + r])) > thresh:  # :26
        j += 1                                          # :27
        if j > min(n, m):
            raise IndexError("BoundsError: adaptive rangefinder exceeded min(m, n) columns")
        Yj = Yfull[:, j - 1]
        Q = Qfull[:, :j - 1]
        QtYj = Q.T @ Yj                                 # :30
        Yj = Yj - Q @ QtYj                              # :31 (rebinding: Yfull is NOT modified)
        Qfull[:, j - 1] += (1 / np.linalg.norm(Yj)) * Yj  # :33-34
        Q = Qfull[:, :j]
        omega = omegas[:, j - 1]                        # :36
        Aomega = A @ omega                              # :37
        QtAomega = Q.T @ Aomega                         # :38
        ynew = Aomega - Q @ QtAomega                    # :39
        Yfull[:, r + j - 1] = ynew                      # :40
        Qj = Qfull[:, j - 1]
        for i in range(j + 1, j + r):                   # :42  (1-based i = j+1 : j+r-1)
            Yi = Yfull[:, i - 1]
            Yi -= np.dot(Qj, Yi) * Qj                   # :44
    return Qfull[:, :j].copy()                          # :47


def rangefinder_adaptive_blocked(A, omegas, epsilon=1e-8, block=16):
    """NO REFERENCE COUNTERPART: CPU statement of the product's opt-in blocked adaptive range
    finder (gsi_rangefinder_adaptive_blocked) -- Halko et al. Alg 4.2 with the
    random vectors consumed `block` at a time: Y = A Omega_b, block Gram-Schmidt against the basis
    (twice), the reference's stopping estimator (src/RandMatFact.jl:26) on the block's fresh
    probes, Householder QR of the block, one more projection + QR of the orthonormalised block.
    Test infrastructure for that mode only."""
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    thresh = epsilon / np.sqrt(200 / np.pi)
    Q = np.zeros((m, 0))
    j = 0
    while True:
        b = min(block, omegas.shape[1] - j)
        if b <= 0:
            raise IndexError("adaptive rangefinder: maxvec basis vectors did not reach epsilon")
        Y = A @ omegas[:, j:j + b]
        for _ in range(2):
            Y = Y - Q @ (Q.T @ Y)
        if not np.max(colnorms(Y)) > thresh:
            break
        Qb, _ = np.linalg.qr(Y)
        # a block that overshoots the rank has columns of pure rounding noise, whose "directions" are
        # not orthogonal to the basis: project the orthonormalised block once more and re-orthonormalise
        Qb, _ = np.linalg.qr(Qb - Q @ (Q.T @ Qb))
        Q = np.hstack([Q, Qb])
        j += b
    return Q


def eig_nystrom(A, Q):
    """src/RandMatFact.jl:92-102 (Halko et al. Alg 5.5). Returns (U, Sigmavec)."""
    B1 = A @ Q                                          # :93
    B2 = Q.T @ B1                                       # :94
    # cholesky(Hermitian(B2)).U : Hermitian() reads the upper triangle -> dpotrf('U')
    c, info = lapack.dpotrf(np.asfortranarray(np.triu(B2)), lower=0)
    if info != 0:
        raise PosDefException(info)
    C = np.triu(c)
    Cinv, info = lapack.dtrtri(np.asfortranarray(C), lower=0)  # inv(::UpperTriangular)
    if info != 0:
        raise SingularException(info)
    F = B1 @ Cinv                                       # :96
    U, Sigmavec, _ = _scipy_svd(F, full_matrices=False, lapack_driver="gesdd")  # :97
    return U, Sigmavec
