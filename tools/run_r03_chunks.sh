mkdir -p gpurun_out
run() { echo "== SWEEP=$1 OPTIONS=$2"; GSI_SWEEP="$1" GSI_OPTIONS="$2" timeout 300 python tools/kcov_probe.py --reps 3 2>&1 | tail -1; }
(
run "64,256,0,0,6" ""
run "64,256,0,0,6" "kcov.chunks=8"
run "64,-1,0,4,6" ""
run "64,-1,0,4,6" "kcov.chunks=4"
run "64,-1,0,4,6" "kcov.chunks=8"
run "64,-1,0,4,6" "kcov.chunks=32"
run "1,0,0,4,6" "kcov.chunks=8"
run "64,-1,0,16,6" "kcov.chunks=8"
) | tee gpurun_out/r03_chunks.log
for c in c1 c2; do timeout 300 python tools/small_probe.py $c > gpurun_out/r03d_probe_$c.json 2> gpurun_out/r03d_probe_$c.err || echo "probe $c failed"; cat gpurun_out/r03d_probe_$c.json; done
timeout 900 python -m pytest tests/test_gpu_blocks.py -q -m gpu -k svd > gpurun_out/r03_t3.log 2>&1; echo "t3 rc=$?"; tail -12 gpurun_out/r03_t3.log
