set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
python tools/ab_small.py > gpurun_out/r2_base_ab_small.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:trace_kernel -c 1 -f"
$NCU -o gpurun_out/r2_base_c2_ch13 python tools/prof_c2.py ch13 > gpurun_out/r2_ncu_c2.log 2>&1
$NCU -o gpurun_out/r2_base_c5_n16 python tools/prof_sweep.py 16 16 960 > gpurun_out/r2_ncu_n16.log 2>&1
$NCU -o gpurun_out/r2_base_c5_n128 python tools/prof_sweep.py 128 16 960 > gpurun_out/r2_ncu_n128.log 2>&1
cat gpurun_out/r2_base_ab_small.log
