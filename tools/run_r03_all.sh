mkdir -p gpurun_out
python tools/svd_accuracy.py > gpurun_out/r03_svd_drivers.json 2> gpurun_out/r03_svd_drivers.err; cat gpurun_out/r03_svd_drivers.json | python -c "
import json,sys
for r in json.load(sys.stdin): print(r['l'], {k:(round(v['ms'],3), round(v['sv_err_over_s1_eps']), round(v['orth_err_eps']), v['sweeps']) for k,v in r.items() if k!='l'})"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r03_gpu_all.log 2>&1; echo "all rc=$?"; tail -15 gpurun_out/r03_gpu_all.log
