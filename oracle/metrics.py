"""Parity metrics.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def singvals_from_Z(Z, K):
    """Z = V * sqrt(S)  =>  S_i = ||Z[:, i]||^2."""
    return np.sum(np.asarray(Z)[:, :K] ** 2, axis=0)


def subspace_sine(Z1, Z2, K):
    """|| (I - Q1 Q1') Q2 ||_2 with Q from QR of the first K columns.
    (acos of sigma_min is ill-conditioned near 0 -- SURVEY.md measurement trap.)"""
    Q1, _ = np.linalg.qr(np.asarray(Z1)[:, :K])
    Q2, _ = np.linalg.qr(np.asarray(Z2)[:, :K])
    R = Q2 - Q1 @ (Q1.T @ Q2)
    return float(np.linalg.norm(R, 2))


def compare_Z(Z, Zref, K):
    """Returns dict(sv_rel, sine, col_up_to_sign, tail_zero)."""
    Z = np.asarray(Z)
    Zref = np.asarray(Zref)
    s = singvals_from_Z(Z, K)
    sref = singvals_from_Z(Zref, K)
    sv_rel = float(np.max(np.abs(s - sref) / sref))
    sine = subspace_sine(Z, Zref, K)
    colerr = 0.0
    for i in range(K):
        nrm = np.linalg.norm(Zref[:, i])
        e = min(np.linalg.norm(Z[:, i] - Zref[:, i]), np.linalg.norm(Z[:, i] + Zref[:, i]))
        colerr = max(colerr, e / nrm)
    tail_zero = bool(np.all(Z[:, K:] == 0.0))
    return dict(sv_rel=sv_rel, sine=sine, col_up_to_sign=float(colerr), tail_zero=tail_zero)
