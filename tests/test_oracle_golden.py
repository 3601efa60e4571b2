"""Pin the CPU oracle to the reference's own golden vectors and unit-test known answers
(SURVEY.md §4.2/§4.3, §8c).  Runs without a GPU."""
import ctypes as C
import math

import numpy as np
import pytest

import rtzlib as R


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


# ---------------------------------------------------------------- end-to-end goldens
def test_chapter14_byte_exact(orc, tmp_path):
    """reference src/main.zig:41-55 with build.zig:62-66 (400 px, 10 spp, seed 0xdeadbeef)."""
    prng, sp, n = R.final_scene(0xDEADBEEF)
    assert n == 485
    cam = R.main_camera(400, 10)
    assert (cam.width, cam.height) == (400, 225)
    rgb = np.zeros((225, 400, 3), np.uint8)
    st = R.Stats()
    assert orc.orc_render_reference(C.byref(cam), sp, n, prng, _u8(rgb), None, C.byref(st)) == 0
    out = tmp_path / "chapter14.ppm"
    assert orc.orc_write_ppm(str(out).encode(), 400, 225, _u8(rgb)) == 0
    assert out.read_bytes() == (R.GOLDEN / "chapter14.ppm").read_bytes()
    # work model figures quoted in SURVEY.md / BASELINE.md
    assert st.samples == 900_000
    assert abs(st.segments / st.samples - 2.644) < 1e-3
    assert st.sphere_tests == st.segments * 485


def test_scene_object_count_abadcafe(orc):
    """reference src/Scene.zig:189-205: 1 + 3 + 22*22 - 3 objects for seed 0xabadcafe."""
    prng, sp, n = R.final_scene(0xABADCAFE)
    assert n == 1 + 3 + 22 * 22 - 3
    assert sp[0].radius == 1000 and list(sp[0].center) == [0, -1000, 0]
    assert [sp[n - 3].mat_type, sp[n - 2].mat_type, sp[n - 1].mat_type] == [R.MAT_DIELECTRIC, R.MAT_LAMBERTIAN, R.MAT_METAL]
    kinds = [sp[i].mat_type for i in range(n)]
    assert kinds.count(R.MAT_LAMBERTIAN) + kinds.count(R.MAT_METAL) + kinds.count(R.MAT_DIELECTRIC) == n


@pytest.mark.parametrize("mode,name,spheres", [
    (R.MODE_LEGACY_SKY, "chapter4", []),
    (R.MODE_LEGACY_FLAT, "chapter5", [((0, 0, -1), 0.5)]),
    (R.MODE_LEGACY_NORMAL, "chapter6", [((0, 0, -1), 0.5), ((0, -100.5, -1), 100)]),
])
def test_deterministic_goldens(orc, mode, name, spheres):
    cam = R.Camera()
    orc.orc_camera_legacy(400, 16.0 / 9.0, mode, C.byref(cam))
    sp = R.sphere_array([R.make_sphere(c, r, 0) for c, r in spheres]) if spheres else (R.Sphere * 1)()
    rgb = np.zeros((225, 400, 3), np.uint8)
    assert orc.orc_render_legacy(C.byref(cam), sp, len(spheres), _u8(rgb), None, None) == 0
    assert rgb.tobytes() == R.read_ppm(R.GOLDEN / f"{name}.ppm")[2]


def test_ppm_binary_writer(orc, tmp_path):
    """reference src/ppm.zig:92-106: 1x1 black -> 15 bytes == test-files/test-binary.ppm."""
    out = tmp_path / "t.ppm"
    px = np.zeros(3, np.uint8)
    assert orc.orc_write_ppm(str(out).encode(), 1, 1, _u8(px)) == 0
    assert out.read_bytes() == (R.GOLDEN / "test-binary.ppm").read_bytes() == b"P6\n1 1\n255\n\x00\x00\x00\n"
    assert orc.orc_write_ppm_ascii(str(out).encode(), 1, 1, _u8(px)) == 0
    assert out.read_bytes() == b"P3\n1 1\n255\n0 0 0\n"  # reference src/ppm.zig:72-90


# ---------------------------------------------------------------- camera known answers
def test_camera_builder_kats(orc):
    """reference src/camera.zig:516-528 (exact f64 equality) and :373-391, :352-369."""
    cam = R.build_camera(400, 16.0 / 9.0, (0, 0, 0), (0, 0, -1), 90)
    assert (cam.width, cam.height) == (400, 225)
    assert list(cam.du) == [8.888888888888888e-2, 0.0, 0.0]
    assert list(cam.dv) == [0.0, -8.888888888888888e-2, 0.0]
    assert list(cam.pixel0) == [-1.773333333333333e1, 9.955555555555554e0, -1e1]
    assert cam.samples_per_pixel == 100 and cam.bounce_max == 50 and cam.pixel_samples_scale == 1.0 / 100
    assert cam.t_min == 1e-3 and math.isinf(cam.t_max)
    assert list(cam.defocus_disk_u) == [0, 0, 0] and cam.defocus_angle == 0
    vw, vh = C.c_double(), C.c_double()
    orc.orc_viewport(400, 225, 90.0, 2.0, C.byref(vw), C.byref(vh))
    height = 2.0 * math.tan(math.radians(90) / 2.0) * 2.0
    assert vh.value == height and vw.value == height * (400 / 225)
    assert orc.orc_image_height(1, 2.0) == 1          # clamps to 1
    assert orc.orc_image_height(400, 1.0) == 400
    for w, h in [(400, 225), (1200, 675), (1920, 1080), (3840, 2160)]:   # Q12
        assert orc.orc_image_height(w, 16.0 / 9.0) == h


def test_setviewport_uses_focus_dist_so_far(orc):
    """Q13: main.zig sets focusDist before setViewport; swapping the order changes the viewport."""
    a = R.build_camera(400, 16 / 9, (13, 2, 3), (0, 0, 0), 20, viewport_focus=10.0, focus=10.0, defocus=0.6)
    b = R.build_camera(400, 16 / 9, (13, 2, 3), (0, 0, 0), 20, viewport_focus=10.0, focus=5.0, defocus=0.6)
    assert list(a.du) == list(b.du) and list(a.pixel0) != list(b.pixel0)


# ---------------------------------------------------------------- geometry KATs
def test_sphere_hit_kats(orc):
    """reference src/sphere.zig:76-136, src/hittable.zig:121-142,185-209."""
    s = R.make_sphere((0, 0, -2), 1, 0)
    h = R.Hit()
    orc.orc_sphere_hit(C.byref(s), R.d3((0, 0, 0)), R.d3((0, 0, -1)), 0.0, 3.0, C.byref(h))
    assert h.hit == 1 and h.t == 1 and list(h.point) == [0, 0, -1] and list(h.normal) == [0, 0, 1] and h.front == 1
    orc.orc_sphere_hit(C.byref(s), R.d3((0, 0, 0)), R.d3((0, 0, -1)), 0.0, 0.0, C.byref(h))
    assert h.hit == 0
    orc.orc_sphere_hit(C.byref(s), R.d3((0, 0, 0)), R.d3((0, 0, 1)), 0.0, 3.0, C.byref(h))
    assert h.hit == 0
    four = R.sphere_array([R.make_sphere((0, 0, -z), 1, 0) for z in (2, 3, 4, 5)])
    orc.orc_list_hit(four, 4, R.d3((0, 0, 0)), R.d3((0, 0, -1)), -6.0, 6.0, C.byref(h))
    assert h.hit == 1 and h.t == 1 and h.index == 0 and list(h.normal) == [0, 0, 1]
    assert R.make_sphere((0, 0, 0), -3, 0).radius == 0  # Sphere.init clamps (Q16)


def test_material_kats(orc):
    """reference src/material.zig:168-281."""
    rec = R.Hit(hit=1, index=0, front=1, t=1.0, point=R.d3((0, 0, -1)), normal=R.d3((0, 0, 1)))
    sc = R.Scatter()
    # lambertian: scattered = (point, normal + randomUnitVec(same seed)), attenuation = albedo
    lam = R.make_sphere((0, 0, -2), 1, R.MAT_LAMBERTIAN, albedo=(0.1, 0.2, 0.5))
    p1, p2 = orc.orc_prng_new(0xABADCAFE), orc.orc_prng_new(0xABADCAFE)
    orc.orc_scatter(C.byref(lam), R.d3((0, 0, 0)), R.d3((0, 0, -1)), C.byref(rec), p1, C.byref(sc))
    u = R.D3()
    orc.orc_random_unit_vec(p2, u)
    assert sc.scattered == 1 and list(sc.origin) == [0, 0, -1]
    exp = [0 + u[0], 0 + u[1], 1 + u[2]]
    if all(x < 1e-8 for x in exp):
        exp = [0, 0, 1]
    assert list(sc.direction) == exp and list(sc.attenuation) == [0.1, 0.2, 0.5]
    # metal fuzz 0: direction == reflect(dir, normal)
    metal = R.make_sphere((0, 0, -2), 1, R.MAT_METAL, albedo=(0.8, 0.8, 0.8), fuzz=0.0)
    orc.orc_scatter(C.byref(metal), R.d3((0, 0, 0)), R.d3((0, 0, -1)), C.byref(rec), p1, C.byref(sc))
    refl = R.D3()
    orc.orc_vec_reflect(R.d3((0, 0, -1)), R.d3((0, 0, 1)), refl)
    assert sc.scattered == 1 and list(sc.direction) == list(refl) == [0, 0, 1]
    # the unit-vector draw happens even with fuzz 0 (Q5): the two streams are now out of step by one vector
    d1 = orc.orc_prng_draws(p1)
    assert d1 > orc.orc_prng_draws(p2)
    # dielectric 1.5 head-on with seed 0xabadcafe: refracts (src/material.zig:222-246)
    glass = R.make_sphere((0, 0, -2), 1, R.MAT_DIELECTRIC, ior=1.5)
    p3 = orc.orc_prng_new(0xABADCAFE)
    orc.orc_scatter(C.byref(glass), R.d3((0, 0, 0)), R.d3((0, 0, -1)), C.byref(rec), p3, C.byref(sc))
    refr = R.D3()
    orc.orc_vec_refract(R.d3((0, 0, -1)), R.d3((0, 0, 1)), 1.0 / 1.5, refr)
    assert list(sc.direction) == list(refr) and list(sc.attenuation) == [1, 1, 1]
    assert orc.orc_prng_draws(p3) == 1
    # total internal reflection consumes NO draw (Q6): inside glass, grazing
    p4 = orc.orc_prng_new(1)
    rec_in = R.Hit(hit=1, index=0, front=0, t=1.0, point=R.d3((0, 0, -1)), normal=R.d3((0, 0, 1)))
    orc.orc_scatter(C.byref(glass), R.d3((0, 0, 0)), R.d3((1, 0, -0.1)), C.byref(rec_in), p4, C.byref(sc))
    assert orc.orc_prng_draws(p4) == 0 and sc.direction[2] > 0
    assert abs(orc.orc_reflectance(1.0, 1.5) - 0.04) < 1e-15
    for p in (p1, p2, p3, p4):
        orc.orc_prng_free(p)


def test_vec_interval_color_kats(orc):
    """reference src/vec.zig:154-162,274-295; src/interval.zig:84-154; src/color.zig:131-135,157-172."""
    assert orc.orc_vec_near_zero(R.d3((0, 0, 0))) == 1
    assert orc.orc_vec_near_zero(R.d3((1, 1, 1))) == 0
    assert orc.orc_vec_near_zero(R.d3((1e-9, 1e-9, 1e-9))) == 1
    assert orc.orc_vec_near_zero(R.d3((-1, -2, -3))) == 1   # no abs (Q1)
    u = R.D3()
    orc.orc_vec_unit(R.d3((1, 0, 2)), u)
    inv = 1.0 / math.sqrt(5.0)
    assert list(u) == [1 * inv, 0.0, 2 * inv]               # multiply by reciprocal (Q2)
    c = R.D3()
    orc.orc_vec_cross(R.d3((1, 0, 0)), R.d3((0, 1, 0)), c)
    assert list(c) == [0, 0, 1]
    assert orc.orc_vec_dot(R.d3((1, 2, 3)), R.d3((4, 5, 6))) == 32 and orc.orc_vec_len(R.d3((3, 4, 0))) == 5
    assert orc.orc_interval_surrounds(0, 1, 0) == 0 and orc.orc_interval_surrounds(0, 1, 0.5) == 1
    assert orc.orc_interval_contains(0, 1, 0) == 1 and orc.orc_interval_contains(0, 1, 1.5) == 0
    assert [orc.orc_interval_clamp(0, 1, x) for x in (-1, 0.5, 2)] == [0, 0.5, 1]
    lin = np.array([[0, .5, .75], [1, 0, 1]], np.float64)
    out = np.zeros((2, 3), np.uint8)
    orc.orc_to_rgb(lin.ctypes.data_as(C.POINTER(C.c_double)), 2, _u8(out))
    assert out.tolist() == [[0, 181, 221], [255, 0, 255]]
    assert [orc.orc_linear_to_gamma(x) for x in (-1.0, 0.0, 4.0)] == [0, 0, 2]


# ---------------------------------------------------------------- RNG (Zig std restated; Philox)
def test_zig_rng_restatement(orc):
    """Xoshiro256++ seeded by SplitMix64: the published SplitMix64 vector for seed 0, and the
    Random.float(f64) construction; the chapter14 byte match above certifies the combination."""
    p = orc.orc_prng_new(0)
    st = (C.c_uint64 * 4)()
    orc.orc_prng_state(p, st)
    # SplitMix64(0) first four outputs (published reference values)
    assert list(st) == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F, 0xF88BB8A8724C81EC]
    s = list(st)
    rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & (2**64 - 1)
    first = (rotl((s[0] + s[3]) & (2**64 - 1), 23) + s[0]) & (2**64 - 1)
    p2 = orc.orc_prng_new(0)
    assert orc.orc_prng_next(p2) == first
    # float(f64): 52 mantissa bits of the word, exponent 1022 - clz(top 12 bits)
    lz = 64 - first.bit_length()
    if lz < 12:
        bits = ((1022 - lz) << 52) | (first & 0xFFFFFFFFFFFFF)
        assert orc.orc_prng_float(p) == np.array([bits], np.uint64).view(np.float64)[0]
    vals = [orc.orc_prng_float(p) for _ in range(20000)]
    assert 0 <= min(vals) and max(vals) < 1 and abs(np.mean(vals) - 0.5) < 0.01
    orc.orc_prng_free(p), orc.orc_prng_free(p2)


def test_philox_known_answers(orc):
    """Random123 kat_vectors for philox4x32-10."""
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in kats:
        out = (C.c_uint32 * 4)()
        orc.orc_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == exp
