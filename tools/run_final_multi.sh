# Multi-GPU evidence of a round (gpurun --gpus N): the sharded parity tests, then the bench line at N ranks.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/final_gpu_dist.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/final_gpu_dist.log; fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_bench_c3_g$N.json 2> gpurun_out/final_bench_c3_g$N.err; echo "bench rc=$?"
tail -c 400 gpurun_out/final_bench_c3_g$N.err; tail -c 3000 gpurun_out/final_bench_c3_g$N.json
