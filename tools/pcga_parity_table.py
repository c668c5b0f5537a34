#!/usr/bin/env python
"""PCGA / RGA parity table: for the cases of the reference's
test/testrpcga.jl:104-138 and the BASELINE configurations C2 / C4 at full size, measured on the
GPU box against the oracle with IDENTICAL xis and forward-model evaluations:

  itn GPU / oracle and istop of the first iteration's LSQR at the package's DEFAULT tolerances,
  relerr of one pcgalsqr iteration at default tolerances and with LSQR run to convergence,
  relerr of one pcgadirect iteration next to eps * cond(bigA), retained rank,
  relerr of the final estimates (default full runs) GPU vs oracle and each vs ground truth,
  bit-identity of the paramstorun batch.

    python tools/pcga_parity_table.py > profiles/r02/pcga_parity_table.json
The assertions of tests/test_gpu_pcga.py / test_gpu_configs.py are set to <= 10x these values."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsi_b200 as gsi            # noqa: E402
import oracle                     # noqa: E402  (checker)
import pcga_cases as pc           # noqa: E402
from gsi_b200.pcga import pcgadirectiteration, pcgalsqriteration, LinearForwardModel  # noqa: E402

EPS = float(np.finfo(float).eps)
rows = []


def direct_rows(name, forward, s0, X, xis, R, y, truth):
    s1 = pcgadirectiteration(forward, s0, X, xis, R, y, pc.DELTA, lambda s, o: None)
    s1o = oracle.pcgadirectiteration(forward, s0, X, xis, R, y, pc.DELTA, lambda s, o: None)
    bigA = oracle.pcgadirect_system(forward, s0, X, xis, R, y, pc.DELTA)[0]
    sv = np.linalg.svd(bigA, compute_uv=False)
    kept = sv[sv > EPS * len(sv) * sv[0]]
    sg = gsi.pcgadirect(forward, s0, X, xis, R, y)
    so = oracle.pcgadirect(forward, s0, X, xis, R, y)
    rows.append({"case": name, "solver": "pcgadirect", "iter1_relerr": pc.relerr(s1, s1o),
                 "eps_cond": EPS * kept[0] / kept[-1], "rank_kept": int(len(kept)), "size": int(len(sv)),
                 "final_relerr_gpu_vs_oracle": pc.relerr(sg, so), "final_gpu_vs_truth": pc.relerr(sg, truth),
                 "final_oracle_vs_truth": pc.relerr(so, truth)})


def lsqr_rows(name, forward, s0, X, xis, R, y, truth, conv_maxiter=None, fm_device=None):
    itg, ito, isg, iso, xrel = pc.lsqr_first_iteration_info(gsi, forward, s0, X, xis, R, y)
    s1 = pcgalsqriteration(forward, s0, X, xis, R, y, pc.DELTA)
    s1o = oracle.pcgalsqriteration(forward, s0, X, xis, R, y, pc.DELTA)
    conv = dict(pc.TIGHT)
    if conv_maxiter:
        conv["maxiter"] = conv_maxiter
    s1t = pcgalsqriteration(forward, s0, X, xis, R, y, pc.DELTA, lsqr_kwargs=conv)
    s1ot = oracle.pcgalsqriteration(forward, s0, X, xis, R, y, pc.DELTA, lsqr_kwargs=conv)
    sg = gsi.pcgalsqr(forward, s0, X, xis, R, y)
    so = oracle.pcgalsqr(forward, s0, X, xis, R, y)
    row = {"case": name, "solver": "pcgalsqr", "lsqr_itn_gpu": itg, "lsqr_itn_oracle": ito, "lsqr_istop_gpu": isg,
           "lsqr_istop_oracle": iso, "lsqr_x_relerr_default": xrel, "iter1_relerr_default": pc.relerr(s1, s1o),
           "iter1_relerr_converged_lsqr": pc.relerr(s1t, s1ot),
           "final_relerr_gpu_vs_oracle": pc.relerr(sg, so), "final_gpu_vs_truth": pc.relerr(sg, truth),
           "final_oracle_vs_truth": pc.relerr(so, truth),
           "paramstorun_bit_identical": pc.paramstorun_bit_identical(gsi, s0, X, xis)}
    if fm_device is not None:
        sd = gsi.pcgalsqr(fm_device, s0, X, xis, R, y)
        row["final_relerr_device_forward_batch_vs_oracle"] = pc.relerr(sd, so)
        row["final_device_forward_batch_vs_truth"] = pc.relerr(sd, truth)
    rows.append(row)


t0 = time.time()
for log2N, log2M, mu in pc.SIMPLE_CASES:
    c = pc.simple_case(log2N, log2M, mu)
    xis = gsi.getxis(c["Q"], c["K"], c["p"], Omega=c["Omega"])
    direct_rows(c["name"], c["forward"], c["s0"], c["X"], xis, c["R"], c["y"], c["truth"])
    if c["lsqr_ok"]:
        lsqr_rows(c["name"], c["forward"], c["s0"], c["X"], xis, c["R"], c["y"], c["truth"])

# rga (test/testrpcga.jl:133-138): default pcgadirect and the F5 case pcgafunc = pcgalsqr
M, N, Nred, mu = 8, 1024, 512, 10.0
rng = np.random.default_rng(N)
forward, p0, X, Q, Omega, R, yobs, truep, pp = pc.setupsimpletest(rng, M, N, mu)
xis = gsi.getxis(Q, M, pp, Omega=Omega)
S = rng.standard_normal((Nred, N)) * (1 / np.sqrt(N))
for nm, fg, fo in (("pcgadirect", gsi.pcgadirect, oracle.pcgadirect), ("pcgalsqr", gsi.pcgalsqr, oracle.pcgalsqr)):
    pg = gsi.rga(forward, p0, X, xis, R, yobs, S, pcgafunc=fg)
    po = oracle.rga(forward, p0, X, xis, R, yobs, S, pcgafunc=fo)
    rows.append({"case": "rga N=1024 M=8 Nred=512", "solver": "rga/" + nm, "final_relerr_gpu_vs_oracle": pc.relerr(pg, po),
                 "final_gpu_vs_truth": pc.relerr(pg, truep), "final_oracle_vs_truth": pc.relerr(po, truep)})

# C2 at BASELINE size
c = pc.config2(full=True)
C = oracle.kernel_cov_dense(0, c["coords"], c["ell"])
op = gsi.GridKernelCovMatrix("exponential", c["grid"], c["ell"])
xis = gsi.getxis(op, c["K"], c["p"], c["q"], Omega=c["Omega"])
xis_ref = oracle.getxis(C, c["Omega"], c["K"], c["p"], c["q"])
xpar = max(min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b) for a, b in zip(xis, xis_ref))
del C
truth, y = pc.config2_truth(c, xis)
H = c["H"]
lsqr_rows(c["name"], lambda s: H @ s, c["s0"], c["X"], xis, c["R"], y, truth, conv_maxiter=20000,
          fm_device=LinearForwardModel(H))
rows[-1]["xis_relerr_up_to_sign_vs_oracle"] = xpar

# C4 at BASELINE size
c = pc.config4(full=True)
lr = gsi.LowRankCovMatrix(c["fields"])
xis = gsi.getxis(lr, c["K"], c["p"], c["q"], Omega=c["Omega"])
lro = pc.GemmLowRankCov(c["fields"])
xis_ref = oracle.getxis(lro, c["Omega"], c["K"], c["p"], c["q"])
xpar = max(min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b) for a, b in zip(xis, xis_ref))
truth, y = pc.config4_truth(c, xis)
pg = gsi.rga(c["forward"], c["s0"], c["X"], xis, c["R"], y, c["S"], pcgafunc=gsi.pcgalsqr)
Sy, SRS = c["S"] @ y, (c["S"] * c["R"][None, :]) @ c["S"].T
po = oracle.pcgalsqr(lambda x: c["S"] @ c["forward"](x), c["s0"], c["X"], xis, SRS, Sy)
rows.append({"case": c["name"], "solver": "rga/pcgalsqr (oracle: pcgalsqr on the sketched triple)",
             "xis_relerr_up_to_sign_vs_oracle": xpar, "final_relerr_gpu_vs_oracle": pc.relerr(pg, po),
             "final_gpu_vs_truth": pc.relerr(pg, truth), "final_oracle_vs_truth": pc.relerr(po, truth)})
print(json.dumps({"seconds": time.time() - t0, "rows": rows}, indent=1))
