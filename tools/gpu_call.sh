for i in 1 2; do for v in 0 6 7; do echo "variant $v: $(RTZ_VARIANT=$v python tools/prof_run.py 500 3 1200)"; done; done
for v in 0 6 7; do echo "variant $v: $(RTZ_VARIANT=$v python tools/ab_small.py | tr '\n' ';')"; done
RTZ_VARIANT=6 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mirror" 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | wc -l
