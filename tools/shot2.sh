#!/bin/bash
# Second GPU call: probe of multi-front schedules, DRAM bytes of the best ones, and one
# `ncu --set full` capture of the product kernel under the picked schedule.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python tools/sweep_probe.py --set fronts --reps 2 --control --out gpurun_out/sweep_probe2.json > gpurun_out/sweep_probe2.log 2>&1
echo "probe rc=$?"; tail -14 gpurun_out/sweep_probe2.log
PICK=$(python tools/pick_schedule.py gpurun_out/sweep_probe2.json 2>/dev/null)
SW=$(echo "$PICK" | sed -n 1p); IDX=$(echo "$PICK" | sed -n 2p)
[ -z "$SW" ] && SW="64,256,0,0,6"; [ -z "$IDX" ] && IDX="0 1 2"
echo "picked GSI_SWEEP=$SW ; DRAM capture of indices: $IDX"
timeout 50 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:kcov --csv --log-file gpurun_out/sweep_probe2_ncu.csv \
    python tools/sweep_probe.py --set fronts --only $IDX --reps 1 --no-warm --out gpurun_out/sweep_probe2_under_ncu.json \
    > gpurun_out/sweep_probe2_ncu.log 2>&1
echo "ncu dram rc=$?"; grep -c kcov gpurun_out/sweep_probe2_ncu.csv
FULLIDX=$(echo $IDX | cut -d" " -f2)
timeout 70 ncu --set full --import-source on --clock-control none -k regex:kcov -c 1 -f -o gpurun_out/prof_kcov_c3_picked \
    python tools/sweep_probe.py --set fronts --only $FULLIDX --reps 1 --no-warm --out gpurun_out/sweep_probe2_full.json \
    > gpurun_out/sweep_probe2_full.log 2>&1
echo "ncu full rc=$? (schedule index $FULLIDX)"; ls -la gpurun_out/prof_kcov_c3_picked.ncu-rep 2>/dev/null
