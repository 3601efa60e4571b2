python tools/run_configs.py > gpurun_out/r2c_configs.log 2>&1; tail -2 gpurun_out/r2c_configs.log | cut -c1-200
