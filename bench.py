#!/usr/bin/env python
"""Headline benchmark: randsvd (K=200, p=10, q=2) of a matrix-free covariance operator.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c5|c1|small] [--impl reference]

One "step" = one complete randsvd(A, K, p, q) of the workload (sketch GEMM, 2q power
products with pivot-faithful LU normalisation, final Householder QR, projection, TSQR +
Jacobi SVD, Z = V sqrt(S)).  Default workload c3 = BASELINE.json configs[2]: 3-D
64x56x56 = 200 704-point Gaussian covariance, single B200; with N > 1 the rows of the
operator and of every iterate are sharded over the ranks (strong scaling, same problem).

metric  = randsvd throughput in FP64 TFLOP/s, F = (2q+2)*2*n^2*(K+p) algorithmic flops
           / wall time of the step (factorisations are in the time, not
          in the numerator);  ms_per_step = randsvd time.
value   : Omega already resident in HBM, Z left in HBM.
e2e     : through the public API with HOST buffers -- Omega uploaded from pinned host
          memory and Z downloaded to pinned host memory inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, grid, ell, K, p, q, description)
    "c3": ("gaussian", (64, 56, 56), (9.0, 7.0, 5.0), 200, 10, 2,
           "BASELINE configs[2]: randsvd of matrix-free 200704-point (64x56x56) 3-D Gaussian covariance, K=200 p=10 q=2"),
    "c5": ("exponential", (1000, 1000), (120.0, 80.0), 200, 10, 2,
           "BASELINE configs[4]: row-sharded randsvd of matrix-free 10^6-point (1000x1000) exponential covariance, K=200 p=10 q=2"),
    "c5s": ("exponential", (512, 500), (60.0, 40.0), 200, 10, 2,
            "reduced configs[4]: 256000-point (512x500) exponential covariance, K=200 p=10 q=2"),
    "small": ("gaussian", (28, 26, 24), (9.0, 7.0, 5.0), 200, 10, 2,
              "reduced configs[2]: 17472-point (28x26x24) 3-D Gaussian covariance, K=200 p=10 q=2"),
}
# dense sample the CPU baseline / reference arm runs (a 200704^2 dense matrix is 322 GB)
CPU_SAMPLE = {"c3": (28, 26, 24), "small": (20, 18, 16), "c5": (132, 132), "c5s": (132, 132)}


def grid_coords(shape):
    axes = [np.arange(s, dtype=np.float64) for s in shape]
    grids = np.meshgrid(*axes, indexing="ij")
    return np.stack([g.ravel(order="F") for g in grids], axis=0)


def randsvd_flops(n, l, q):
    return (2 * q + 2) * 2.0 * n * n * l


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_fp64_peak():
    """cuBLAS DGEMM 8192^3 best-of-5 (MEASURED_PEAKS.json carries no FP64 entry)."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best * 1e-9


def cpu_oracle_run(workload, steps, warmup):
    """Times the oracle (reference algorithm restated on SciPy/OpenBLAS, all host cores) on a
    bounded dense sample of the workload.  Returns (tflops, ms_per_step, sample, cores)."""
    import oracle
    kind, grid, ell, K, p, q, _ = WORKLOADS[workload]
    sgrid = CPU_SAMPLE[workload]
    kid = {"exponential": 0, "gaussian": 1, "powerlaw": 2}[kind]
    coords = grid_coords(sgrid)
    n = coords.shape[1]
    C = oracle.kernel_cov_dense(kid, coords, ell[:len(sgrid)])
    Omega = np.random.default_rng(0).standard_normal((n, K + p))
    for _ in range(warmup):
        oracle.randsvd(C, Omega, K, p, q)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.randsvd(C, Omega, K, p, q)
    dt = (time.perf_counter() - t0) / steps
    sample = (f"dense n={n} ({'x'.join(map(str, sgrid))} grid, {kind}) oracle randsvd K={K} p={p} q={q}; "
              f"flops-normalised (the full-size dense matrix cannot be materialised)")
    return randsvd_flops(n, K + p, q) / dt * 1e-12, dt * 1e3, sample, os.cpu_count()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="gsi", choices=["gsi", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--generation", default="table", choices=["table", "arithmetic"],
                    help="kernel values: lattice-table look-up (structured grid) or exp/sqrt arithmetic from coordinates")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind, grid, ell, K, p, q, desc = WORKLOADS[args.workload]
    l = K + p

    if args.impl == "reference":
        # The reference (pure Julia) cannot run here: its CPU algorithm restated on
        # SciPy/OpenBLAS is timed on the host cores; rank 0 only.
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        tf, ms, sample, cores = cpu_oracle_run(args.workload, steps, 1)
        print(json.dumps({
            "impl": "reference", "metric": "randsvd_fp64_tflops", "value": tf, "unit": "TFLOP/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "sample": sample},
            "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import gsi_b200 as gsi
    torch.cuda.set_device(local_rank)
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(gsi.Context.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().tolist())
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"

    peak = measured_fp64_peak() if rank == 0 else None

    ctx = gsi.Context(local_rank, rank, world, uid)
    gsi.set_default_context(ctx)
    coords = grid_coords(grid)
    n = coords.shape[1]
    row0, mloc = gsi.partition_rows(n, world, rank)
    if args.generation == "table":
        op = gsi.GridKernelCovMatrix(kind, grid, ell, ctx=ctx, row0=row0, mloc=mloc)
    else:
        op = gsi.KernelCovMatrix(kind, coords, ell, ctx=ctx, row0=row0, mloc=mloc)

    # host-seeded Omega in pinned memory (the reference draws randn(n, l) on the host)
    omega_pinned = torch.empty((l, n), dtype=torch.float64, pin_memory=True)      # (l, n) C-order == (n, l) F-order
    Omega_h = omega_pinned.numpy().T
    Omega_h[...] = np.random.default_rng(0).standard_normal((n, l))
    z_pinned = torch.empty((l, mloc), dtype=torch.float64, pin_memory=True)
    Z_h = z_pinned.numpy().T
    Omega_d = gsi.DeviceMatrix.from_host(ctx, Omega_h)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        Z = gsi.randsvd(op, K, p, q, Omega=Omega_d, device_out=True)
        Z.free()

    def step_e2e():
        Z = gsi.randsvd(op, K, p, q, Omega=Omega_h, device_out=True)     # uploads Omega from pinned host memory
        Z.numpy(out=Z_h)                                                 # downloads this rank's rows of Z
        Z.free()

    def timed(fn, steps, with_timing):
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ctx.launch_count(reset=True)
        if with_timing:
            ctx.gemm_timing(enable=True)
            ctx.phase_timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count()
        gemm = ctx.gemm_timing(enable=False) if with_timing else None
        phases = ctx.phase_timing() if with_timing else None
        clocks = sampler.stop() if rank == 0 else None
        if dist is not None:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
            lt = torch.tensor([launches], dtype=torch.float64, device="cuda")
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
        return ms, launches, gemm, clocks, phases

    for _ in range(args.warmup):
        step_resident()
    ms, launches, gemm, clocks, phases = timed(step_resident, args.steps, True)
    ms_per_step = ms / args.steps
    F = randsvd_flops(n, l, q)
    value = F / (ms_per_step * 1e-3) * 1e-12

    e2e = None
    if not args.no_e2e:
        step_e2e()
        ms2, _, _, _, _ = timed(step_e2e, args.steps, False)
        e2e_ms = ms2 / args.steps
        e2e = {"value": F / (e2e_ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(world * n * l * 8), "d2h_bytes_per_step": int(n * l * 8)}

    # k-sweep schedule of the product kernel (gsi_ctx_set_option / GSI_SWEEP): groups,div,hint,window,epoch_shift
    schedule = ",".join(str(ctx.get_option(k)) for k in ("kcov.sweep_groups", "kcov.sweep_div", "kcov.l2_hint",
                                                         "kcov.window", "kcov.epoch_shift"))
    if rank == 0:
        gemm_ms, gemm_launches, gemm_flops = gemm
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if (world == 1 and args.generation == "table" and args.workload in tj
                    and tj[args.workload].get("schedule") == schedule):
                traffic = tj[args.workload]["bytes_per_launch"]      # from the committed ncu capture of this schedule
        except Exception:
            traffic = None
        ach = gemm_flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else None
        roof = {"bound": "tensor", "kernel": "kcov_gemm_kernel (matrix-free covariance x tall-skinny, DMMA.8x8x4)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic,
                "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) best of 5 measured in this run; "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "launches_timed": gemm_launches,
                "algorithmic_flops_per_launch": gemm_flops / max(gemm_launches, 1),
                "gemm_share_of_step": gemm_ms / ms}
        out = {"metric": "randsvd_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": desc, "n": n, "K": K, "p": p, "q": q, "kernel": kind, "ell": list(ell),
                          "normaliser": "LU_REF", "parallelism": f"row-shard x{world}",
                          "kernel_values": ("lattice table look-up (structured grid, n distinct values)"
                                            if args.generation == "table" else "exp/sqrt arithmetic from coordinates"),
                          "l2": "operand streams (X 366 MB/pass) exceed L2; no flush needed",
                          "kcov_schedule": schedule, "svd_fused": ctx.get_option("svd.fused")},
               "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
               "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()}}
        if not args.no_cpu_baseline and world == 1:
            tf, cms, sample, cores = cpu_oracle_run(args.workload, 1, 1)
            out["cpu_baseline"] = {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                                   "ms_per_step_sample": cms}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
