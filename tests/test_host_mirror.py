"""The C++ host mirror (raytracing-with-zig_b200/host) against the oracle: the host-side f64 work
(scene generation with the Zig-exact Xoshiro stream, CameraBuilder maths) must be bit-identical,
because it feeds the GPU path.  CPU-only except the last test."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import rtzlib as R


@pytest.fixture(scope="module")
def host(pkg):
    return importlib.import_module("raytracing-with-zig_b200.host_api")


@pytest.mark.parametrize("seed", [0xDEADBEEF, 0xABADCAFE, 0, 12345])
def test_generate_world_bit_identical_to_oracle(host, orc, seed):
    buf, n = host.generate_world(seed)
    prng, obuf, on = R.final_scene(seed)
    assert n == on
    if seed in (0xDEADBEEF, 0xABADCAFE):
        assert n == 485  # reference src/Scene.zig:204
    assert bytes(buf)[: n * C.sizeof(R.Sphere)] == bytes(obuf)[: n * C.sizeof(R.Sphere)]


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_generate_sweep_identical_to_oracle(host, orc, n):
    """BASELINE config 5 scenes: product host generator == oracle generator, byte for byte."""
    buf, got = host.generate_sweep(0xDEADBEEF, n)
    prng = orc.orc_prng_new(0xDEADBEEF)
    obuf = (R.Sphere * n)()
    assert orc.orc_generate_sweep(prng, n, obuf) == n == got
    assert bytes(buf) == bytes(obuf)
    assert buf[0].radius == 1000 and [buf[i].radius for i in (1, 2, 3)] == [1, 1, 1]


def test_generate_chapter13_identical(host, orc):
    buf, n = host.generate_chapter13()
    obuf, on = R.chapter13_scene()
    assert n == on == 5 and bytes(buf) == bytes(obuf)


def test_camera_builder_identical_and_kats(host, orc):
    cam = host.camera_build(400, 16.0 / 9.0, (0, 0, 0), (0, 0, -1), 90)
    assert list(cam.du) == [8.888888888888888e-2, 0.0, 0.0]              # reference src/camera.zig:520-528
    assert list(cam.dv) == [0.0, -8.888888888888888e-2, 0.0]
    assert list(cam.pixel0) == [-1.773333333333333e1, 9.955555555555554e0, -1e1]
    assert (cam.width, cam.height, cam.samples_per_pixel, cam.bounce_max) == (400, 225, 100, 50)
    for width, spp, seed in [(400, 10, 0xDEADBEEF), (1200, 500, 7), (3840, 2000, None)]:
        a = host.main_camera(width, spp, seed)
        b = R.main_camera(width, spp, seed)
        assert bytes(a) == bytes(b)
    a = host.camera_build(400, 16 / 9, (-2, 2, 1), (0, 0, -1), 20, focus_dist=3.4, defocus_angle=10.0, spp=100, seed=1)
    b = R.build_camera(400, 16 / 9, (-2, 2, 1), (0, 0, -1), 20, viewport_focus=3.4, focus=3.4, defocus=10.0, spp=100, seed=1)
    assert bytes(a) == bytes(b)
    # lookFrom == lookAt: unit() of a zero vector panics in the reference -> error, not NaNs
    with pytest.raises(Exception):
        host.camera_build(400, 16 / 9, (1, 1, 1), (1, 1, 1), 20)


@pytest.mark.gpu
def test_main_reproduces_mirror_image_and_file_format(host, orc, tmp_path):
    """main() end to end on the GPU: images/<fileName> is a P6 file whose pixels equal the mirror's."""
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with pytest.raises(Exception):
            host.run_main(400, 10, "x.ppm", 0xDEADBEEF)   # images/ must exist (Q18) -> RTZ_ERR_IO
        os.mkdir("images")
        st = host.run_main(400, 10, "chapter14.ppm", 0xDEADBEEF)
    finally:
        os.chdir(cwd)
    w, h, body, raw = R.read_ppm(tmp_path / "images" / "chapter14.ppm")
    assert (w, h) == (400, 225) and raw.endswith(b"\n") and len(raw) == 15 + 270000 + 1
    prng, sp, n = R.final_scene(0xDEADBEEF)
    cam = R.main_camera(400, 10, seed=0xDEADBEEF)
    exp = np.zeros((225 * 400, 3), np.uint8)
    mst = R.Stats()
    assert orc.orc_render_mirror(C.byref(cam), sp, n, 0xDEADBEEF, 8, None, exp.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 None, C.byref(mst)) == 0
    assert body == exp.tobytes()
    assert st.samples == 900000 and st.segments == mst.segments
    # same frame as the reference golden up to sampling noise (different RNG stream by design)
    gold = np.frombuffer(R.read_ppm(R.GOLDEN / "chapter14.ppm")[2], np.uint8).astype(float)
    rmse = np.sqrt(np.mean((gold - np.frombuffer(body, np.uint8)) ** 2))
    assert rmse < 1.3 * 13.04, rmse   # seed-to-seed noise floor at 10 spp


@pytest.mark.gpu
def test_host_hittable_list_hit_runs_on_device(host):
    four = R.sphere_array([R.make_sphere((0, 0, -z), 1, 0) for z in (2, 3, 4, 5)])
    pkg = importlib.import_module("raytracing-with-zig_b200")
    h = pkg.binding.rtz_hit()
    rc = host.hostlib().rtzh_list_hit(C.cast(four, C.POINTER(pkg.rtz_sphere)), 4, R.d3((0, 0, 0)), R.d3((0, 0, -1)), -6.0,
                                      6.0, C.byref(h))
    assert rc == 0 and h.hit == 1 and h.t == 1.0 and list(h.normal) == [0, 0, 1]


# ------------------------------------------------------------------ P3 writer and Color converters
def test_color_from_value_and_from_rgb_are_the_reference_formulas(host):
    """Color.fromValue / Color.fromRgb (reference src/color.zig:30-38, 53-61): channel / 255.999 in f64.  Host-only."""
    assert host.color_from_value((255 << 16) | (0 << 8) | 255) == (255 / 255.999, 0.0, 255 / 255.999)
    assert host.color_from_value(0x123456) == (0x12 / 255.999, 0x34 / 255.999, 0x56 / 255.999)
    assert host.color_from_rgb(255, 0, 255) == (255 / 255.999, 0.0, 255 / 255.999)
    assert host.color_from_rgb(1, 2, 3) == (1 / 255.999, 2 / 255.999, 3 / 255.999)


@pytest.mark.gpu
def test_color_converters_kats(host):
    """The reference's own KATs, src/color.zig:122-163 — Color.toRgb (gamma 2, clamp, trunc(256 x)) runs on the device:
    fromValue(0xFF00FF).toRgb() == (255, 0, 255); Color(1, 0, 1).toValue() == 0xFF00FF; fromRgb(255, 0, 255).toRgb()
    == (255, 0, 255); Color(0, .5, .75).toRgb() == (0, 181, 221)."""
    assert host.color_to_rgb(host.color_from_value(0xFF00FF)) == (255, 0, 255)
    assert host.color_to_value((1.0, 0.0, 1.0)) == 0xFF00FF
    assert host.color_to_rgb(host.color_from_rgb(255, 0, 255)) == (255, 0, 255)
    assert host.color_to_rgb((0.0, 0.5, 0.75)) == (0, 181, 221)
    assert host.color_to_rgb((-1.0, 0.0, 4.0)) == (0, 0, 255)          # linearToGamma: -1 -> 0, 4 -> 2 -> clamp .999


@pytest.mark.gpu
def test_ppm_save_ascii_and_binary_kats(host, tmp_path):
    """PPM.save (ASCII P3, src/ppm.zig:25-39) and PPM.saveBinary (:42-60) of the product's host mirror against the
    reference's KATs: a fresh 1x1 PPM saves as "P3\\n1 1\\n255\\n0 0 0\\n" (:72-90) and as test-files/test-binary.ppm
    (:92-106); a filled image writes one "r g b" line per pixel, row-major."""
    p3, p6 = tmp_path / "test.ppm", tmp_path / "test-binary.ppm"
    host.ppm_save(p3, 1, 1)
    assert p3.read_bytes() == b"P3\n1 1\n255\n0 0 0\n"
    host.ppm_save(p6, 1, 1, binary=True)
    assert p6.read_bytes() == (R.GOLDEN / "test-binary.ppm").read_bytes()
    host.ppm_save(p3, 2, 2, [(0.0, 0.5, 0.75), (1.0, 0.0, 1.0), (0.25, 0.25, 0.25), (4.0, -1.0, 0.0)])
    assert p3.read_bytes() == b"P3\n2 2\n255\n0 181 221\n255 0 255\n128 128 128\n255 0 0\n"
    host.ppm_save(p6, 2, 2, [(0.0, 0.5, 0.75), (1.0, 0.0, 1.0), (0.25, 0.25, 0.25), (4.0, -1.0, 0.0)], binary=True)
    assert p6.read_bytes() == b"P6\n2 2\n255\n" + bytes([0, 181, 221, 255, 0, 255, 128, 128, 128, 255, 0, 0]) + b"\n"
    with pytest.raises(Exception):
        host.ppm_save(tmp_path / "nodir" / "x.ppm", 1, 1)


@pytest.mark.gpu
def test_main_with_num_gpus_option(host, tmp_path):
    """-DnumGpus=0 (every GPU of the box): main() goes through rtz_render_multi from this one process and writes
    the same file as the single-GPU run."""
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        os.mkdir("images")
        st1 = host.run_main(160, 6, "one.ppm", 0xDEADBEEF)
        stn = host.run_main(160, 6, "all.ppm", 0xDEADBEEF, num_gpus=0)
    finally:
        os.chdir(cwd)
    import torch
    assert stn.gpus == torch.cuda.device_count() and st1.gpus == 1
    assert (tmp_path / "images" / "one.ppm").read_bytes() == (tmp_path / "images" / "all.ppm").read_bytes()
    assert (stn.samples, stn.segments) == (st1.samples, st1.segments)
