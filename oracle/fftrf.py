"""Oracle restatement of FFTRF.powerlaw_structuredgrid (reference:
src/FFTRF.jl:40-100) -- used only to GENERATE synthetic input fields for the
LowRankCovMatrix / rga configurations.  TEST INFRASTRUCTURE ONLY.

`phi` replaces `randn(size(S))` (src/FFTRF.jl:75).
"""
import numpy as np


def _fouriercoords(N):
    # vcat(collect(0:N), -1 * collect((N-1):-1:1))           src/FFTRF.jl:88
    return np.concatenate([np.arange(0, N + 1), -np.arange(N - 1, 0, -1)]).astype(np.float64)


def powerlaw_structuredgrid(Ns, k0, dk, beta, rng=None, phi=None):
    Ns = list(Ns)
    d = len(Ns)
    fc = [_fouriercoords(N) for N in Ns]
    # computesqrtS_f: S_f has Julia size (2Ns[2], 2Ns[1][, 2Ns[3]]) with linear index j
    # decoding  fc1 index = div(j-1, len2) (2-D), fc2 index = rem(j-1, len2)   :45-49
    if d == 2:
        f2, f1 = np.meshgrid(fc[1], fc[0], indexing="ij")       # array [i2, i1], i2 fastest
        S = f1 ** 2 + f2 ** 2
    elif d == 3:
        f2, f1, f3 = np.meshgrid(fc[1], fc[0], fc[2], indexing="ij")
        S = f1 ** 2 + f2 ** 2 + f3 ** 2
    else:
        raise ValueError("unsupported dimension")
    with np.errstate(divide="ignore"):
        S = S ** (0.25 * beta)                                   # :64
    S[np.isinf(S)] = 0.0                                         # :65-67
    if phi is None:
        phi = rng.standard_normal(S.shape)                       # :75
    assert phi.shape == S.shape
    result = S * (np.cos(2 * np.pi * phi) + 1j * np.sin(2 * np.pi * phi))  # :77-79 cospi/sinpi
    kc = np.fft.ifftn(result)                                    # :92
    # reducek :9-38 -- finalk[j, i(, h)] = real(k[i, j(, h)]) over the first halves
    if d == 2:
        finalk = np.real(kc[:kc.shape[0] // 2, :kc.shape[1] // 2]).T.copy()
    else:
        finalk = np.real(kc[:kc.shape[0] // 2, :kc.shape[1] // 2, :kc.shape[2] // 2]).transpose(1, 0, 2).copy()
    std = np.std(finalk, ddof=1)                                 # Statistics.std (corrected)
    mean = np.mean(finalk)
    return dk * (finalk - mean) / std + k0                       # :96-98
