"""Edge cases and error behaviour of the C-ABI path on the GPU (shapes that stress the
tiling: single row/column, non-multiples of the 64-row / 8-column / 32-point tiles, widest
iterate; dimension mismatches; repeated calls / buffer reuse from the pool)."""
import numpy as np
import pytest

import oracle
from gpu_util import gsi, relerr  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,l", [(1, 1), (2, 1), (31, 3), (33, 9), (63, 17), (65, 47), (127, 1), (129, 256), (1000, 211)])
def test_kernelcov_ragged_shapes(gsi, n, l):
    rng = np.random.default_rng(n * 7 + l)
    coords = rng.uniform(0, 20, size=(2, n))
    ell = [4.0, 3.0]
    C = oracle.kernel_cov_dense(0, coords, ell, sigma2=2.0, nugget=0.5)
    X = rng.standard_normal((n, l))
    Y = gsi.KernelCovMatrix("exponential", coords, ell, sigma2=2.0, nugget=0.5) @ X
    assert Y.shape == (n, l)
    assert relerr(Y, C @ X) < 1e-12


@pytest.mark.parametrize("m,n,l", [(1, 1, 1), (3, 5, 2), (64, 33, 8), (31, 64, 9), (130, 70, 255)])
def test_dense_ragged_shapes(gsi, m, n, l):
    rng = np.random.default_rng(m + 3 * n + l)
    A = rng.standard_normal((m, n))
    X, Xt = rng.standard_normal((n, l)), rng.standard_normal((m, l))
    op = gsi.DenseMatrix(A)
    assert relerr(op @ X, A @ X) < 1e-13
    assert relerr(op.T @ Xt, A.T @ Xt) < 1e-13


def test_wide_host_matrix_is_chunked(gsi):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((300, 300))
    X = rng.standard_normal((300, 700))
    assert relerr(gsi.DenseMatrix(A) @ X, A @ X) < 1e-13


def test_dimension_mismatch_and_bad_arguments(gsi):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((50, 40))
    op = gsi.DenseMatrix(A)
    with pytest.raises(gsi.DimensionMismatch):
        op @ rng.standard_normal((41, 3))
    with pytest.raises(gsi.DimensionMismatch):
        gsi.randsvd(op, 5, 2, 1, Omega=rng.standard_normal((39, 7)))
    with pytest.raises(gsi.GsiError):
        gsi.randsvd(op, 60, 10, 1)            # l > min(size(A))
    with pytest.raises(gsi.GsiError):
        gsi.KernelCovMatrix("gaussian", rng.standard_normal((2, 10)), [1.0, -1.0])
    with pytest.raises(KeyError):
        gsi.KernelCovMatrix("matern", rng.standard_normal((2, 10)), [1.0, 1.0])
    with pytest.raises(gsi.GsiError):
        gsi.KernelCovMatrix("gaussian", rng.standard_normal((4, 10)), [1.0] * 4)   # d > 3


def test_l_equals_one_and_q_zero(gsi):
    rng = np.random.default_rng(2)
    A = rng.standard_normal((40, 1)) @ rng.standard_normal((1, 40))
    Omega = rng.standard_normal((40, 1))
    for q in (0, 1, 2):
        Z = gsi.randsvd(A, 1, 0, q, Omega=Omega)
        Zr = oracle.randsvd(A, Omega, 1, 0, q)
        c = oracle.compare_Z(Z, Zr, 1)
        assert c["sv_rel"] < 1e-10 and c["sine"] < 1e-8


def test_repeated_calls_reuse_pool_and_agree_bitwise(gsi):
    coords = oracle.grid_coords((30, 20))
    op = gsi.KernelCovMatrix("gaussian", coords, [5.0, 4.0])
    Omega = np.random.default_rng(3).standard_normal((600, 25))
    Z1 = gsi.randsvd(op, 20, 5, 2, Omega=Omega)
    Z2 = gsi.randsvd(op, 20, 5, 2, Omega=Omega)
    assert np.array_equal(Z1, Z2)              # deterministic: fixed reduction orders everywhere
    assert np.all(Z1[:, 20:] == 0.0)


def test_device_resident_api(gsi):
    """Omega supplied as a device buffer, Z returned as a device buffer."""
    ctx = gsi.default_context()
    coords = oracle.grid_coords((25, 20))
    op = gsi.KernelCovMatrix("exponential", coords, [6.0, 4.0])
    Omega = np.random.default_rng(4).standard_normal((500, 30))
    Od = gsi.DeviceMatrix.from_host(ctx, Omega)
    Zd, S = gsi.randsvd(op, 25, 5, 2, Omega=Od, device_out=True, return_singular_values=True)
    Z = Zd.numpy()
    assert np.array_equal(Z, gsi.randsvd(op, 25, 5, 2, Omega=Omega))
    assert np.allclose(S[:25], oracle.singvals_from_Z(Z, 25), rtol=1e-12)
    before = ctx.launch_count(reset=True)
    assert before > 0


def test_pinned_host_array_and_large_pageable_upload(gsi):
    """gsi_host_alloc: a page-locked column-major array (plain DMA upload); and a pageable source large enough
    to take the threaded bounce-buffer path with a pitched (row-block) source."""
    ctx = gsi.default_context()
    rng = np.random.default_rng(11)
    P = ctx.pinned_empty((5000, 37))
    assert P.flags.f_contiguous and P.shape == (5000, 37)
    P[...] = rng.standard_normal(P.shape)
    for layout in (gsi.LAYOUT_TALL, gsi.LAYOUT_COLMAJOR):
        d = gsi.DeviceMatrix.from_host(ctx, P, layout)
        assert np.array_equal(d.numpy(), P)
        d.free()
    del P                                          # releases the block (gsi_host_free)
    A = rng.standard_normal((400000, 36))          # 115 MB: more rows than one staging block, slices of 16 MB
    A = np.asfortranarray(A)
    d = gsi.DeviceMatrix.from_host(ctx, A)
    assert np.array_equal(d.numpy(), A)
    d.free()
    B = np.asfortranarray(rng.standard_normal((70000, 30)))
    d = gsi.DeviceMatrix.from_host(ctx, B, gsi.LAYOUT_COLMAJOR)
    assert np.array_equal(d.numpy(), B)
    d.free()
