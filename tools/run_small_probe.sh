set -x
mkdir -p gpurun_out
for c in c1 c2 c2dense sketch; do python tools/small_probe.py $c > gpurun_out/r03_probe_$c.json 2> gpurun_out/r03_probe_$c.err || echo "probe $c failed"; cat gpurun_out/r03_probe_$c.json; done
for c in c1 c2 c2dense; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03_launches_$c.csv python tools/small_probe.py $c 1 > gpurun_out/r03_ncu_$c.log 2>&1
python tools/launch_summary.py gpurun_out/r03_launches_$c.csv 4 > gpurun_out/r03_launches_${c}_summary.txt; head -30 gpurun_out/r03_launches_${c}_summary.txt
done
