"""Oracle restatement of `IterativeSolvers.lsqr(A, b)` as called at
reference src/lsqr.jl:54.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

IterativeSolvers.jl (compat "0.9", reference Project.toml:18) is a third-party
dependency that is NOT vendored under /root/reference and cannot be installed
here (no Julia, no network).  This file restates the published Paige & Saunders
LSQR recurrence (ACM TOMS 8(1), 1982) in the form IterativeSolvers 0.9 uses,
with that package's defaults: x0 = 0, damp = 0, atol = btol = sqrt(eps),
conlim = 1/sqrt(eps), maxiter = max(size(A)).  Stopping tests 1-7 as in the
original LSQR.  PARITY UNPINNED at 1e-8 (pinned only by the reference's own
2e-2 end-to-end tests, test/testrpcga.jl:128-129).

One deliberate fidelity note: IterativeSolvers 0.9 accumulates the condition
estimate as `ddnorm += norm(w/rho)` using the UPDATED w (not norm^2 of the
old w as in the 1982 paper).  That only feeds stopping test 3/6 (Acond); the
variant is selectable with `ddnorm_mode`.
"""
import numpy as np


def lsqr(A, b, damp=0.0, atol=None, btol=None, conlim=None, maxiter=None,
         ddnorm_mode="iterativesolvers", return_info=False):
    b = np.asarray(b, dtype=np.float64)
    m, n = A.shape
    eps = np.finfo(np.float64).eps
    if atol is None:
        atol = np.sqrt(eps)
    if btol is None:
        btol = np.sqrt(eps)
    if conlim is None:
        conlim = 1.0 / np.sqrt(eps)
    if maxiter is None:
        maxiter = max(m, n)
    x = np.zeros(n)
    itn = 0
    istop = 0
    ctol = 1.0 / conlim if conlim > 0 else 0.0
    Anorm = Acond = ddnorm = res2 = xnorm = xxnorm = z = sn2 = 0.0
    cs2 = -1.0
    dampsq = damp * damp

    u = b - A @ x
    v = np.zeros(n)
    beta = np.linalg.norm(u)
    alpha = 0.0
    At = A.T
    if beta > 0:
        u = u * (1.0 / beta)
        v = At @ u
        alpha = np.linalg.norm(v)
    if alpha > 0:
        v = v * (1.0 / alpha)
    w = v.copy()
    Arnorm = alpha * beta
    if Arnorm == 0:
        return (x, dict(itn=0, istop=0)) if return_info else x

    rhobar = alpha
    phibar = bnorm = rnorm = beta
    while itn < maxiter and istop == 0:
        itn += 1
        tmpm = A @ v
        u = -alpha * u + tmpm
        beta = np.linalg.norm(u)
        if beta > 0:
            u = u * (1.0 / beta)
            Anorm = np.sqrt(Anorm * Anorm + alpha * alpha + beta * beta + dampsq)
            tmpn = At @ u
            v = -beta * v + tmpn
            alpha = np.linalg.norm(v)
            if alpha > 0:
                v = v * (1.0 / alpha)

        rhobar1 = np.sqrt(rhobar * rhobar + dampsq)
        cs1 = rhobar / rhobar1
        sn1 = damp / rhobar1
        psi = sn1 * phibar
        phibar = cs1 * phibar

        rho = np.sqrt(rhobar1 * rhobar1 + beta * beta)
        cs = rhobar1 / rho
        sn = beta / rho
        theta = sn * alpha
        rhobar = -cs * alpha
        phi = cs * phibar
        phibar = sn * phibar
        tau = sn * phi

        t1 = phi / rho
        t2 = -theta / rho
        if ddnorm_mode == "paige-saunders":
            dk = w * (1.0 / rho)
            ddnorm += np.dot(dk, dk)
        x = x + t1 * w
        w = t2 * w + v
        if ddnorm_mode == "iterativesolvers":
            wrho = w * (1.0 / rho)
            ddnorm += np.linalg.norm(wrho)

        delta = sn2 * rho
        gambar = -cs2 * rho
        rhs = phi - delta * z
        zbar = rhs / gambar
        xnorm = np.sqrt(xxnorm + zbar * zbar)
        gamma = np.sqrt(gambar * gambar + theta * theta)
        cs2 = gambar / gamma
        sn2 = theta / gamma
        z = rhs / gamma
        xxnorm += z * z

        Acond = Anorm * np.sqrt(ddnorm)
        res1 = phibar * phibar
        res2 = res2 + psi * psi
        rnorm = np.sqrt(res1 + res2)
        Arnorm = alpha * abs(tau)

        test1 = rnorm / bnorm
        test2 = Arnorm / (Anorm * rnorm) if (Anorm * rnorm) != 0 else np.inf
        test3 = 1.0 / Acond if Acond != 0 else np.inf
        t1 = test1 / (1.0 + Anorm * xnorm / bnorm)
        rtol = btol + atol * Anorm * xnorm / bnorm

        if itn >= maxiter:
            istop = 7
        if 1 + test3 <= 1:
            istop = 6
        if 1 + test2 <= 1:
            istop = 5
        if 1 + t1 <= 1:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
    if return_info:
        return x, dict(itn=itn, istop=istop, Anorm=Anorm, Acond=Acond, rnorm=rnorm,
                       Arnorm=Arnorm, xnorm=xnorm)
    return x
