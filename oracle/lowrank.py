"""Oracle restatement of the implicit operators (reference: src/lowrank.jl).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np


class LowRankCovMatrix:
    """src/lowrank.jl:14-30 -- sample covariance S S'/(N-1) of mean-removed fields."""
    __array_ufunc__ = None   # let `ndarray @ self` fall through to __rmatmul__

    def __init__(self, samples):
        samples = [np.asarray(s, dtype=np.float64) for s in samples]
        means = np.zeros(len(samples[0]))
        for s in samples:                               # :19-23
            means += s
        means = means / len(samples)                    # :24
        self.samples = [s - means for s in samples]     # :25-27

    # adjoint/transpose return self (:38-44)
    @property
    def T(self):
        return self

    @property
    def shape(self):                                    # :50-52
        n = len(self.samples[0])
        return (n, n)

    def size(self, i):                                  # :54-60 (1-based i)
        if i == 1 or i == 2:
            return len(self.samples[0])
        raise ValueError(f"there is no {i}-th dimension in a LowRankCovMatrix")

    def mul_vec(self, x):                               # :75-81, :135-139
        v = np.zeros(self.shape[0])
        N = len(self.samples)
        for s in self.samples:
            v += (1.0 / (N - 1)) * (s * np.dot(x, s))
        return v

    def __matmul__(self, B):
        B = np.asarray(B, dtype=np.float64)
        if B.ndim == 1:
            return self.mul_vec(B)
        result = np.zeros((self.shape[0], B.shape[1]))  # :115-121  N x ger!
        N = len(self.samples)
        for s in self.samples:
            result += (1.0 / (N - 1)) * np.outer(s, B.T @ s)
        return result

    def __rmatmul__(self, B):
        """`B * A` (:123-129) and `B' * A = (A * B)'` (:131-133) coincide
        mathematically; the Adjoint method is the one randsvd hits (Q' * A)."""
        B = np.asarray(B, dtype=np.float64)
        return (self @ B.T).T

    def dense(self):
        n = self.shape[0]
        return self @ np.eye(n)


class PCGALowRankMatrix:
    """src/lowrank.jl:32-36 -- [HQH'+R, HX; HX', 0] with HQH' = sum eta_i eta_i'."""
    __array_ufunc__ = None

    def __init__(self, etas, HX, R):
        self.etas = [np.asarray(e, dtype=np.float64) for e in etas]
        self.HX = np.asarray(HX, dtype=np.float64)
        self.R = R

    @property
    def T(self):                                        # :38-44
        return self

    @property
    def shape(self):                                    # :62-65
        s = len(self.etas[0]) + 1
        return (s, s)

    def size(self, i):                                  # :67-73
        if i == 1 or i == 2:
            return len(self.etas[0]) + 1
        raise ValueError(f"there is no {i}-th dimension in a PCGALowRankMatrix")

    def _Rmul(self, x):
        R = self.R
        if np.isscalar(R):
            return R * x
        if hasattr(R, "ndim") and R.ndim == 1:          # diagonal stored as a vector
            return R * x
        return R @ x

    def mul(self, x):                                   # :83-97
        x = np.asarray(x, dtype=np.float64)
        v = np.empty(len(x))
        xshort = x[:-1]
        v[:-1] = self._Rmul(xshort)
        v[-1] = np.dot(self.HX, xshort)
        for eta in self.etas:
            dotp = np.dot(eta, xshort)
            v[:-1] += eta * dotp
        v[:-1] += self.HX * x[-1]
        return v

    def __matmul__(self, x):                            # :109-113
        return self.mul(x)

    def dense(self):
        nobs = len(self.etas[0])
        HQH = np.zeros((nobs, nobs))
        for eta in self.etas:
            HQH += np.outer(eta, eta)
        R = self.R
        if np.isscalar(R):
            Rm = R * np.eye(nobs)
        elif hasattr(R, "ndim") and R.ndim == 1:
            Rm = np.diag(R)
        else:
            Rm = np.asarray(R.todense()) if hasattr(R, "todense") else np.asarray(R)
        top = np.hstack([HQH + Rm, self.HX[:, None]])
        bot = np.hstack([self.HX[None, :], np.zeros((1, 1))])
        return np.vstack([top, bot])
