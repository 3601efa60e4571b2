# scratch: the command list of the last `gpurun -- 'bash tools/gpu_call.sh'` call (development aid)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3
