"""CPU oracle for the GeostatInversion.jl randomized low-rank hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product path
(`geostatinversion.jl_b200/`) may import this package.  It is used by
`tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s CPU-baseline /
`--impl reference` arms, and only as the checker / reported baseline.

What it is: a line-by-line NumPy/SciPy restatement of the reference's Julia
code (`/root/reference/src/*.jl`, cited per function) that calls the same
LAPACK routines Julia's LinearAlgebra calls (dgemm, dgetrf, dgeqp3, dorgqr,
dgesdd, dpotrf, dtrtri).  The random matrix Omega is an explicit argument
(Julia's `randn` stream is not reproduced; the product takes a host-seeded
Omega too, so both sides consume identical inputs).

PARITY PINNING STATUS
  * The reference ships NO golden vectors / stored outputs.
  * Julia is not installed in this image, so the reference cannot be run to
    generate vectors.  Against Julia's own floating-point output this oracle
    is therefore "parity unpinned".
  * It IS pinned against every property test and the one closed-form
    known-answer test the reference holds for this path
    (test/testrmf.jl:13-18,21-29; test/testrpcga.jl:10-58,83-131), restated in
    tests/test_oracle_*.py, and against a second independent implementation
    of the LU-with-unpermuted-L normaliser (pure-Python GEPP in
    oracle/gepp_ref.py) so that the pivot rule (first maximal |value|,
    LAPACK idamax) is checked rather than assumed.
  * IterativeSolvers.lsqr (v0.9, not vendored in /root/reference) is restated
    from the published Paige & Saunders recurrence with that package's
    defaults: at 1e-8 "parity unpinned" (only pinned by the reference's 2e-2
    end-to-end tests).
"""
from .randmatfact import (colnorms, rangefinder_adaptive, rangefinder_adaptive_blocked, rangefinder_fixed,
                          randsvd, eig_nystrom, lu_L_unpermuted)
from .lowrank import LowRankCovMatrix, PCGALowRankMatrix
from .lsqr import lsqr
from .pcga import (pcgalsqr, pcgalsqriteration, pcgadirect, pcgadirectiteration, pcgadirect_system, pinv, rga,
                   getxis)
from .kernels import kernel_cov_dense, scaled_coords, grid_coords
from .metrics import singvals_from_Z, subspace_sine, compare_Z
