mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_pcga.py tests/test_gpu_configs.py tests/test_gpu_randsvd.py tests/test_golden.py -q -m gpu > gpurun_out/r03_t1.log 2>&1; echo "t1 rc=$?"; tail -25 gpurun_out/r03_t1.log
for c in c1 c2; do timeout 300 python tools/small_probe.py $c > gpurun_out/r03c_probe_$c.json 2> gpurun_out/r03c_probe_$c.err || echo "probe $c failed"; cat gpurun_out/r03c_probe_$c.json; done
for c in c1 c2; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03c_launches_$c.csv python tools/small_probe.py $c 1 > gpurun_out/r03c_ncu_$c.log 2>&1
python tools/launch_summary.py gpurun_out/r03c_launches_$c.csv 4 > gpurun_out/r03c_launches_${c}_summary.txt; head -14 gpurun_out/r03c_launches_${c}_summary.txt
done
