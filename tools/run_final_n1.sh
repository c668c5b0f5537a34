# Final single-GPU evidence of a round: the default bench line (what the driver runs), then the ncu launch list of the
# same command (kernel shares; numbers printed under ncu are never bench values).
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_c3_g1.json 2> gpurun_out/final_bench_c3_g1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/final_bench_c3_g1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final_launches_c3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-parity > gpurun_out/final_ncu_list.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/final_launches_c3.csv 3 > gpurun_out/final_launches_c3_summary.txt; head -24 gpurun_out/final_launches_c3_summary.txt
python bench.py --workload dense --steps 5 --warmup 3 > gpurun_out/final_bench_dense.json 2> gpurun_out/final_bench_dense.err; echo "dense rc=$?"
