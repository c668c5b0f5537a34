mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_panel.py tests/test_gpu_blocks.py tests/test_gpu_randsvd.py tests/test_golden.py -x -q -m gpu > gpurun_out/r03_t1.log 2>&1; echo "t1 rc=$?"; tail -5 gpurun_out/r03_t1.log
for c in c1 c2; do timeout 300 python tools/small_probe.py $c > gpurun_out/r03i_probe_$c.json 2> gpurun_out/r03i_probe_$c.err || echo "probe $c failed"; cat gpurun_out/r03i_probe_$c.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03i_launches_c2.csv python tools/small_probe.py c2 1 > gpurun_out/r03i_ncu_c2.log 2>&1
python tools/launch_summary.py gpurun_out/r03i_launches_c2.csv 4 > gpurun_out/r03i_launches_c2_summary.txt; head -7 gpurun_out/r03i_launches_c2_summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r03i_bench_c3.json 2> gpurun_out/r03i_bench_c3.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03i_bench_c3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['achieved'], d['gpu_launches'], d['parity'])
PY
