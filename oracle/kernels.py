"""Oracle for the matrix-free covariance-kernel operator (NEW operator: the
reference has no kernel function at all; the definition below is
the one fixed in include/gsi_b200.h).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

    u_i      = x_i ./ ell                      (coordinates scaled once)
    r2(i,j)  = sum_k (u_i[k] - u_j[k])^2       (k = 0..d-1 in order, fma-free)
    EXPONENTIAL : k = exp(-sqrt(r2))
    GAUSSIAN    : k = exp(-0.5 * r2)
    POWERLAW    : k = (1 + r2)^(-beta) = exp(-beta * log1p(r2))
    C[i,j]   = sigma2 * k + nugget * (i == j)
"""
import numpy as np

EXPONENTIAL, GAUSSIAN, POWERLAW = 0, 1, 2


def grid_coords(shape, spacing=None):
    """Structured-grid point coordinates, d x n, first axis fastest (Julia
    column-major linear index of an array of size `shape`)."""
    d = len(shape)
    spacing = [1.0] * d if spacing is None else list(spacing)
    axes = [np.arange(s, dtype=np.float64) * h for s, h in zip(shape, spacing)]
    grids = np.meshgrid(*axes, indexing="ij")
    return np.stack([g.ravel(order="F") for g in grids], axis=0)


def scaled_coords(coords, ell):
    coords = np.asarray(coords, dtype=np.float64)
    ell = np.asarray(ell, dtype=np.float64).reshape(-1, 1)
    return coords / ell


def kernel_cov_dense(kind, coords, ell, sigma2=1.0, nugget=0.0, beta=1.0,
                     rows=None, cols=None):
    """Materialise C[rows, cols] (all by default)."""
    u = scaled_coords(coords, ell)
    d, n = u.shape
    ui = u if rows is None else u[:, rows]
    uj = u if cols is None else u[:, cols]
    r2 = np.zeros((ui.shape[1], uj.shape[1]))
    for k in range(d):
        diff = ui[k][:, None] - uj[k][None, :]
        r2 += diff * diff
    if kind == EXPONENTIAL:
        C = np.exp(-np.sqrt(r2))
    elif kind == GAUSSIAN:
        C = np.exp(-0.5 * r2)
    elif kind == POWERLAW:
        C = np.exp(-beta * np.log1p(r2))
    else:
        raise ValueError("unknown kernel kind")
    C *= sigma2
    if nugget != 0.0:
        ri = np.arange(n) if rows is None else np.asarray(rows)
        cj = np.arange(n) if cols is None else np.asarray(cols)
        C[ri[:, None] == cj[None, :]] += nugget
    return C
