N=$(nvidia-smi -L | wc -l)
python bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}_c3.json 2> gpurun_out/r2_bench_n${N}_c3.err; tail -2 gpurun_out/r2_bench_n${N}_c3.err; cat gpurun_out/r2_bench_n${N}_c3.json
python bench.py --gpus $N --steps 2 --warmup 1 --workload c4 > gpurun_out/r2_bench_n${N}_c4.json 2> gpurun_out/r2_bench_n${N}_c4.err; tail -2 gpurun_out/r2_bench_n${N}_c4.err; cat gpurun_out/r2_bench_n${N}_c4.json
