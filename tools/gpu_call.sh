for i in 1 2 3; do
for L in tools/ab/librtz_base.so tools/ab/librtz_b64.so; do echo "== $L: $(RTZ_LIB=$L python tools/prof_run.py 500 2 1200)"; done; done
for L in tools/ab/librtz_base.so tools/ab/librtz_b64.so; do echo "== $L"; RTZ_LIB=$L python tools/wave_ab.py "11" "300,400,512" | cut -c1-600; done
RTZ_LIB=tools/ab/librtz_b64.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "final_scene_matches_mirror or independent_of_the_schedule or edge_cases or interleaved" 2>&1 | tail -2
