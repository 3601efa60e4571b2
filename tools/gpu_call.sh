python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_final_scene_400spp_vs_both_reference_renders 2>&1 | tail -5
