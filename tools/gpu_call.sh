python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_final_scene_400spp_vs_both_reference_renders 2>&1 | tail -4
python tools/shard_ab.py 8
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"trace_kernel|drain_kernel|resolve_kernel" -c 3 python tools/prof_run.py 500 1 1200 2>&1 | grep -E "rtz::|gpu__time"
python tools/ab_small.py
