mkdir -p gpurun_out
run() { echo "== grid=$1 SWEEP=$2"; GSI_SWEEP="$2" timeout 300 python tools/kcov_probe.py --grid $1 --reps 5 2>&1 | tail -1; }
(
run 32,32,28 "64,256,0,0,6"
run 32,32,28 "1,0,0,0,6"
run 64,32,28 "64,256,0,0,6"
run 64,56,32 "64,256,0,0,6"
run 64,56,56 "64,256,0,0,6"
) | tee gpurun_out/r03_l2fit.log
