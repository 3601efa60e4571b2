python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_final_scene_400spp_vs_both_reference_renders 2>&1 | tail -5
python tools/shard_ab.py 8
python tools/tail_timeline.py
python tools/ab_small.py
for i in 1 2; do echo "r1: $(RTZ_LIB=tools/ab/librtz_r1.so python tools/prof_run.py 500 3 1200)"; echo "cur: $(python tools/prof_run.py 500 3 1200)"; done
