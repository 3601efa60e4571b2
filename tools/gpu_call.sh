python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/wave_ab.py "11,0" "16,32,64,128,256" > gpurun_out/r2_wave_ab6.log 2>&1; cat gpurun_out/r2_wave_ab6.log
