#!/bin/bash
# One GPU call: GPU parity suite, sweep-schedule probe, DRAM bytes of each schedule (ncu),
# and a bench run under the schedule the probe picks.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01b.log 2>&1; PYRC=$?
echo "pytest rc=$PYRC"; tail -4 gpurun_out/pytest_gpu_r01b.log
timeout 90 python tools/sweep_probe.py --reps 2 --out gpurun_out/sweep_probe.json > gpurun_out/sweep_probe.log 2>&1
echo "probe rc=$?"; tail -14 gpurun_out/sweep_probe.log
SW=$(python tools/pick_schedule.py gpurun_out/sweep_probe.json 2>/dev/null || echo "64,256,0,0,6")
FUSED=0; [ "$PYRC" = "0" ] && FUSED=1
echo "picked GSI_SWEEP=$SW GSI_SVD_FUSED=$FUSED"
GSI_SWEEP=$SW GSI_SVD_FUSED=$FUSED timeout 110 python bench.py --warmup 3 --steps 2 --no-cpu-baseline \
    > gpurun_out/bench_c3_picked.json 2> gpurun_out/bench_c3_picked.err
echo "bench rc=$? (GSI_SWEEP=$SW GSI_SVD_FUSED=$FUSED)"; cat gpurun_out/bench_c3_picked.json | cut -c1-600
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:kcov --csv --log-file gpurun_out/sweep_probe_ncu.csv \
    python tools/sweep_probe.py --reps 1 --no-warm --out gpurun_out/sweep_probe_under_ncu.json > gpurun_out/sweep_probe_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/sweep_probe_ncu.log
