# Last single-GPU call of a round: the whole GPU suite, smoke(), the latency-bound configurations.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final_gpu_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/final_gpu_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/final_smoke.log
for c in c1 c2 c2dense sketch; do timeout 300 python tools/small_probe.py $c > gpurun_out/final_probe_$c.json 2> gpurun_out/final_probe_$c.err || echo "probe $c failed"; cat gpurun_out/final_probe_$c.json; done
