#!/bin/bash
# Round-2 opener (needs ~3 GPU-minutes): why is the L2-served X stream 13 % slower than the
# HBM-streamed one?  Two `ncu --set full` captures of ONE C3 product launch --
# default schedule (index 0 of the "coherent" set) and one coherent windowed front (index 1) --
# and the side-by-side digest.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for IDX in 0 1; do
    timeout 150 ncu --set full --import-source on --clock-control none -k regex:kcov -c 1 -f \
        -o gpurun_out/prof_kcov_c3_sched$IDX \
        python tools/sweep_probe.py --set coherent --only $IDX --reps 1 --no-warm \
        --out gpurun_out/sweep_probe_full$IDX.json > gpurun_out/sweep_probe_full$IDX.log 2>&1
    echo "capture $IDX rc=$?"
done
python tools/ncu_stalls.py gpurun_out/prof_kcov_c3_sched0.ncu-rep gpurun_out/prof_kcov_c3_sched1.ncu-rep \
    | tee gpurun_out/kcov_l2_question_digest.txt
# Does the penalty need the lattice-table look-ups?  Same two schedules with arithmetic generation.
timeout 60 python tools/sweep_probe.py --set coherent --only 0 1 --generation arithmetic --reps 2 \
    --out gpurun_out/sweep_probe_arith.json > gpurun_out/sweep_probe_arith.log 2>&1
echo "arith probe rc=$?"; tail -3 gpurun_out/sweep_probe_arith.log
# Is it the burst with which an L2 hit lands in shared memory?  Same two schedules, X tiles in 4 paced chunks
# (run tests/test_gpu_experimental.py::test_kcov_paced_fetch_is_bit_identical first).
timeout 60 python tools/sweep_probe.py --set coherent --only 0 1 --pace 4 --reps 2 \
    --out gpurun_out/sweep_probe_paced.json > gpurun_out/sweep_probe_paced.log 2>&1
echo "paced probe rc=$?"; tail -3 gpurun_out/sweep_probe_paced.log
# The cleanest test (and a possible win: 5 % of the default launch's warp samples wait for X tiles): default
# schedule, but every shared-memory fill made an L2 hit by a bulk prefetch 2 / 4 k-tiles ahead.
for AHEAD in 2 4; do
    timeout 40 python tools/sweep_probe.py --set coherent --only 0 --prefetch $AHEAD --reps 2 \
        --out gpurun_out/sweep_probe_pref$AHEAD.json > gpurun_out/sweep_probe_pref$AHEAD.log 2>&1
    echo "prefetch $AHEAD rc=$?"; tail -1 gpurun_out/sweep_probe_pref$AHEAD.log
done
