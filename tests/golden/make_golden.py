"""Generates the golden fixtures in this directory FROM THE ORACLE (oracle/, the CPU
restatement of the reference).  The reference itself is Julia and cannot run in the build
image, and it ships no stored vectors, so these fixtures pin the oracle's
own output: they guard the oracle against drift (tests/test_golden.py, CPU) and give the GPU
parity tests a committed target that does not depend on the LAPACK build of the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def main():
    out = {}
    # (1) dense exact-rank matrix, testrmf.jl-style construction (makeA), K = rank
    rng = np.random.default_rng(2017)
    A = rng.standard_normal((120, 12)) @ rng.standard_normal((12, 120))
    Om = np.random.default_rng(0).standard_normal((120, 15))
    out["dense_A"], out["dense_Omega"] = A, Om
    out["dense_Z_q2"] = oracle.randsvd(A, Om, 12, 3, 2)
    out["dense_Q_q0"] = oracle.rangefinder_fixed(A, Om, 0)
    # (2) full-rank 2-D exponential covariance (rank > K+p: exercises the unpermuted-L rule)
    coords = oracle.grid_coords((16, 12))
    ell = np.array([5.0, 3.5])
    C = oracle.kernel_cov_dense(0, coords, ell, sigma2=1.3, nugget=0.02)
    Om2 = np.random.default_rng(1).standard_normal((192, 24))
    out["cov_coords"], out["cov_ell"], out["cov_Omega"] = coords, ell, Om2
    out["cov_Z_q3"] = oracle.randsvd(C, Om2, 20, 4, 3)
    out["cov_L"] = oracle.lu_L_unpermuted(C @ Om2)
    # (3) 3-D Gaussian covariance, matrix-free operator definition
    coords3 = oracle.grid_coords((7, 6, 5))
    ell3 = np.array([3.1, 2.7, 2.3])
    C3 = oracle.kernel_cov_dense(1, coords3, ell3)
    Om3 = np.random.default_rng(2).standard_normal((210, 18))
    out["g3_coords"], out["g3_ell"], out["g3_Omega"] = coords3, ell3, Om3
    out["g3_CX"] = C3 @ Om3
    out["g3_Z_q2"] = oracle.randsvd(C3, Om3, 15, 3, 2)
    # (4) PCGA saddle-point operator and LSQR
    rng = np.random.default_rng(3)
    E = rng.standard_normal((30, 8))
    HX = rng.standard_normal(30)
    R = np.full(30, 1e-2)
    b = np.concatenate([rng.standard_normal(30), [0.0]])
    Aop = oracle.PCGALowRankMatrix([E[:, i] for i in range(8)], HX, R)
    x, info = oracle.lsqr(Aop, b, atol=1e-14, btol=1e-14, conlim=1e16, return_info=True)
    out["pcga_E"], out["pcga_HX"], out["pcga_R"], out["pcga_b"] = E, HX, R, b
    out["pcga_Ab"] = Aop @ b
    out["pcga_x_tight"] = x
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
