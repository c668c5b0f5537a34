"""Second, independent statement of the reference's LU normaliser: textbook
unblocked Gaussian elimination with partial pivoting written as plain loops
(LAPACK dgetf2 semantics: pivot = FIRST row of maximal |value| (idamax), row
interchange, multiply by the reciprocal pivot, rank-1 update), returning the
in-place unit-lower-trapezoidal factor *without* undoing the interchanges --
i.e. exactly what Julia's `lu(Y).L` holds (src/RandMatFact.jl:60-61).

TEST INFRASTRUCTURE ONLY.  Used to pin oracle.lu_L_unpermuted (which calls
LAPACK's recursive dgetrf) and as the small-case checker of the CUDA GEPP.
"""
import numpy as np


def gepp_L_unpermuted(Y):
    A = np.array(Y, dtype=np.float64)
    m, n = A.shape
    k = min(m, n)
    piv = np.zeros(k, dtype=np.int64)
    for j in range(k):
        col = np.abs(A[j:, j])
        p = j + int(np.argmax(col))          # np.argmax returns the first maximum
        piv[j] = p
        if A[p, j] == 0.0:
            raise ArithmeticError(f"SingularException({j + 1})")
        if p != j:
            A[[j, p], :] = A[[p, j], :]
        A[j + 1:, j] *= 1.0 / A[j, j]
        if j + 1 < n:
            A[j + 1:, j + 1:] -= np.outer(A[j + 1:, j], A[j, j + 1:])
    L = np.tril(A[:, :k], -1)
    L[np.arange(k), np.arange(k)] = 1.0
    return L, piv
