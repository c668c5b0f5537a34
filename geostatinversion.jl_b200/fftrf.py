"""`FFTRF` surface (reference src/FFTRF.jl) over the B200 library: power-law random fields on a
structured grid, sampled in batches on the device so that `getxis(samplefield, numfields, ...)`
(src/GeostatInversion.jl:29-38) builds its LowRankCovMatrix without a host round trip.

The random phases are drawn on the host (seedable: `randn(size(S))`, src/FFTRF.jl:75) and uploaded.
"""
import ctypes as C
import numpy as np

from ._lib import check, LAYOUT_COLMAJOR
from .core import DeviceMatrix, default_context


def _doubled_shape(Ns):
    """size(S) of src/FFTRF.jl:45,52: (2 Ns[2], 2 Ns[1] [, 2 Ns[3]])."""
    Ns = [int(v) for v in Ns]
    if len(Ns) == 2:
        return (2 * Ns[1], 2 * Ns[0])
    if len(Ns) == 3:
        return (2 * Ns[1], 2 * Ns[0], 2 * Ns[2])
    raise ValueError(f"unsupported dimension: {len(Ns)}")


def sample_fields_device(Ns, k0, dk, beta, numfields=None, phi=None, rng=None, ctx=None):
    """numfields power-law fields as the columns of a COLMAJOR DeviceMatrix (prod(Ns) x numfields).
    phi: optional (numfields, *size(S)) array replacing the `randn(size(S))` of every field."""
    ctx = ctx or default_context()
    shape = _doubled_shape(Ns)
    if phi is None:
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        phi = rng.standard_normal((int(numfields),) + shape)
    phi = np.asarray(phi, dtype=np.float64)
    if phi.shape[1:] != shape:
        raise ValueError(f"phi must have shape (numfields, {shape})")
    nf = phi.shape[0]
    # column f = field f's phases in Julia's (column-major) linear order
    P = np.empty((int(np.prod(shape)), nf), order="F")
    for f in range(nf):
        P[:, f] = phi[f].ravel(order="F")
    Pd = DeviceMatrix.from_host(ctx, P, LAYOUT_COLMAJOR)
    out = DeviceMatrix(ctx, int(np.prod(Ns)), nf, LAYOUT_COLMAJOR)
    dims = (C.c_int64 * len(Ns))(*[int(v) for v in Ns])
    try:
        check(ctx._lib.gsi_fftrf_powerlaw(ctx._h, len(Ns), dims, float(k0), float(dk), float(beta), Pd._h, out._h))
    finally:
        Pd.free()
    return out


def powerlaw_structuredgrid(Ns, k0, dk, beta, phi=None, rng=None, ctx=None):
    """powerlaw_structuredgrid(Ns, k0, dk, beta) -> array of size Ns (src/FFTRF.jl:83-100)."""
    d = sample_fields_device(Ns, k0, dk, beta, 1, None if phi is None else np.asarray(phi)[None], rng, ctx)
    try:
        return d.numpy()[:, 0].reshape([int(v) for v in Ns], order="F")
    finally:
        d.free()


class PowerLawFieldSampler:
    """A `samplefield` for getxis (src/GeostatInversion.jl:29-38, test/testrpcga.jl:87): called with no
    arguments it returns one field as a vector; `sample_device(numfields)` draws a whole batch on the
    device, which `getxis` uses when it sees this type."""

    def __init__(self, Ns, k0, dk, beta, rng=None, ctx=None):
        self.Ns, self.k0, self.dk, self.beta = [int(v) for v in Ns], k0, dk, beta
        self.rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        self.ctx = ctx

    def __call__(self):
        return powerlaw_structuredgrid(self.Ns, self.k0, self.dk, self.beta, rng=self.rng, ctx=self.ctx).ravel(order="F")

    def sample_device(self, numfields):
        return sample_fields_device(self.Ns, self.k0, self.dk, self.beta, numfields, rng=self.rng, ctx=self.ctx)
