python tools/run_configs.py > gpurun_out/r2c_configs.log 2>&1; tail -3 gpurun_out/r2c_configs.log; ls gpurun_out/*.json | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; cat gpurun_out/r2c_bench_n1.json
