"""Statistical link of the parity chain, on the CPU: the f32 device mirror (which the GPU must
match bit for bit, test_gpu_parity.py) against the f64 restatement of the reference (which is
pinned byte for byte by chapter14.ppm, test_oracle_golden.py).

Gates: RMSE(mirror_N, ref_N) <= 1.25 x RMSE(ref_N, ref'_N) (the seed-to-seed
noise floor measured in the same test), |mean signed error| <= 0.25 8-bit levels per channel,
segments/sample within 2 %, depth-cap terminations <= 0.05 % of samples.
"""
import ctypes as C

import numpy as np
import pytest

import rtzlib as R

THREADS = 8


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _ref(orc, cam, sp, n, seed):
    rgb = np.zeros((cam.height, cam.width, 3), np.uint8)
    st = R.Stats()
    assert orc.orc_render_philox64(C.byref(cam), sp, n, seed, THREADS, _u8(rgb), None, C.byref(st)) == 0
    return rgb.astype(np.float64), st


def _mirror(orc, cam, sp, n, seed):
    rgb = np.zeros((cam.height, cam.width, 3), np.uint8)
    st = R.Stats()
    assert orc.orc_render_mirror(C.byref(cam), sp, n, seed, THREADS, None, _u8(rgb), None, C.byref(st)) == 0
    return rgb.astype(np.float64), st


def _rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)))


def _gates(m, mst, a, ast, b):
    floor = _rmse(a, b)
    rm = _rmse(m, a)
    bias = (m - a).mean(axis=(0, 1))
    assert rm <= 1.25 * floor, (rm, floor)
    assert np.abs(bias).max() <= 0.25, bias
    seg_m, seg_a = mst.segments / mst.samples, ast.segments / ast.samples
    assert abs(seg_m - seg_a) / seg_a < 0.02, (seg_m, seg_a)
    assert mst.depth_capped / mst.samples <= 5e-4
    return rm, floor, bias


def test_final_scene_mirror_vs_f64_reference(orc):
    prng, sp, n = R.final_scene(0xDEADBEEF)
    cam = R.main_camera(240, 64)   # 240x135, 64 spp: 2.07 M samples per render
    m, mst = _mirror(orc, cam, sp, n, 0xDEADBEEF)
    a, ast = _ref(orc, cam, sp, n, 1)
    b, _ = _ref(orc, cam, sp, n, 2)
    rm, floor, bias = _gates(m, mst, a, ast, b)
    # the noise floor itself follows the 41/sqrt(N) law measured on the reference
    assert 0.7 * 41 / 8 < floor < 1.4 * 41 / 8, floor
    assert abs(mst.segments / mst.samples - 2.644) / 2.644 < 0.02


@pytest.mark.parametrize("preset", ["ch12", "ch13"])
def test_chapter13_scene_mirror_vs_f64_reference(orc, preset):
    cams = {
        "ch12": dict(look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20),
        "ch13": dict(look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20, defocus=10.0, viewport_focus=3.4, focus=3.4),
    }
    sp, n = R.chapter13_scene()
    cam = R.build_camera(200, 16.0 / 9.0, spp=100, **cams[preset])
    m, mst = _mirror(orc, cam, sp, n, 0xDEADBEEF)
    a, ast = _ref(orc, cam, sp, n, 1)
    b, _ = _ref(orc, cam, sp, n, 2)
    _gates(m, mst, a, ast, b)
    assert abs(ast.segments / ast.samples - 3.584) / 3.584 < 0.03


def test_finite_t_max_and_zero_bounces_mirror_vs_reference(orc):
    """Scene.interval is editable in the reference (src/Scene.zig:21): a finite max and bounceMax = 0
    behave the same in the device mirror and in the f64 restatement."""
    sp, n = R.chapter13_scene()
    cam = R.build_camera(160, 16.0 / 9.0, (-2, 2, 1), (0, 0, -1), 20, spp=64)
    cam.t_max = 0.9
    m, mst = _mirror(orc, cam, sp, n, 5)
    a, ast = _ref(orc, cam, sp, n, 1)
    b, _ = _ref(orc, cam, sp, n, 2)
    _gates(m, mst, a, ast, b)
    cam.t_max = float("inf")
    full, fst = _mirror(orc, cam, sp, n, 5)
    assert mst.segments < fst.segments
    cam.bounce_max = 0
    z, zst = _mirror(orc, cam, sp, n, 5)
    za, zast = _ref(orc, cam, sp, n, 1)
    assert not z.any() and not za.any() and zst.segments == 0 == zast.segments and zst.samples == zast.samples
