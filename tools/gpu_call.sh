N=$(nvidia-smi -L | wc -l)
if [ "$N" = "8" ]; then python -m pytest tests/test_multi_gpu.py tests/test_host_mirror.py -m gpu -v 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed|plugins|collected" > gpurun_out/r2_multi_gpu_tests_n8.log; tail -3 gpurun_out/r2_multi_gpu_tests_n8.log; fi
if [ "$N" = "1" ]; then python __graft_entry__.py smoke 2>&1 | tail -2; fi
LAUNCH="python"; if [ "$N" != "1" ]; then LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
$LAUNCH bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}_c3.json 2> gpurun_out/r2_bench_n${N}_c3.err; tail -2 gpurun_out/r2_bench_n${N}_c3.err; cat gpurun_out/r2_bench_n${N}_c3.json | cut -c1-400
if [ "$N" != "1" ]; then $LAUNCH bench.py --gpus $N --steps 2 --warmup 1 --workload c4 > gpurun_out/r2_bench_n${N}_c4.json 2> gpurun_out/r2_bench_n${N}_c4.err; tail -2 gpurun_out/r2_bench_n${N}_c4.err; cat gpurun_out/r2_bench_n${N}_c4.json | cut -c1-400; fi
