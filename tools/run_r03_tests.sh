mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_pcga.py tests/test_gpu_randsvd.py tests/test_golden.py -x -q -m gpu > gpurun_out/r03_t1.log 2>&1; echo "t1 rc=$?"; tail -5 gpurun_out/r03_t1.log
for c in c1 c2; do timeout 300 python tools/small_probe.py $c > gpurun_out/r03h_probe_$c.json 2> gpurun_out/r03h_probe_$c.err || echo "probe $c failed"; cat gpurun_out/r03h_probe_$c.json; done
python tools/svd_accuracy.py > gpurun_out/r03h_svd_drivers.json 2> gpurun_out/r03h_svd_drivers.err; cat gpurun_out/r03h_svd_drivers.json | python -c "
import json,sys
for r in json.load(sys.stdin): print(r['l'], {k:(round(v['ms'],3), round(v['sv_err_over_s1_eps']), round(v['orth_err_eps']), v['sweeps']) for k,v in r.items() if k!='l'})"
