"""ctypes view of host/librtz_host.so — the C++ mirror of the reference's host API
(Scene.init/generateWorld/generateChapter13, Camera.builder(...).build(), main).  Product code:
it links librtz.so and never the oracle."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import binding as B

HOST_LIB = Path(__file__).resolve().parent / "host" / "librtz_host.so"
_lib = None


def hostlib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not HOST_LIB.exists():
            raise ImportError(f"{HOST_LIB} is missing: run build.py (no fallback)")
        B.lib()  # librtz.so first, so the dependency resolves from the in-tree copy
        l = C.CDLL(str(HOST_LIB))
        u64, i32, f64 = C.c_uint64, C.c_int32, C.c_double
        l.rtzh_scene_generate_world.restype = u64
        l.rtzh_scene_generate_world.argtypes = [u64, i32, C.POINTER(B.rtz_sphere), u64]
        l.rtzh_scene_generate_sweep.restype = u64
        l.rtzh_scene_generate_sweep.argtypes = [u64, u64, C.POINTER(B.rtz_sphere)]
        l.rtzh_scene_generate_chapter13.restype = u64
        l.rtzh_scene_generate_chapter13.argtypes = [C.POINTER(B.rtz_sphere), u64]
        l.rtzh_camera_build.restype = i32
        l.rtzh_camera_build.argtypes = [u64, f64, B.D3, B.D3, B.D3, f64, f64, f64, u64, u64, u64, i32,
                                        C.POINTER(B.rtz_camera)]
        l.rtzh_main.restype = i32
        l.rtzh_main.argtypes = [u64, u64, C.c_char_p, u64, i32, i32, C.POINTER(B.rtz_stats)]
        l.rtzh_ppm_save.restype = i32
        l.rtzh_ppm_save.argtypes = [C.c_char_p, u64, u64, C.POINTER(f64), i32]
        l.rtzh_color_from_value.restype = None
        l.rtzh_color_from_value.argtypes = [C.c_uint32, B.D3]
        l.rtzh_color_from_rgb.restype = None
        l.rtzh_color_from_rgb.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, B.D3]
        l.rtzh_color_to_value.restype = i32
        l.rtzh_color_to_value.argtypes = [B.D3, C.POINTER(C.c_uint32)]
        l.rtzh_color_to_rgb.restype = i32
        l.rtzh_color_to_rgb.argtypes = [B.D3, C.c_uint8 * 3]
        l.rtzh_list_hit.restype = i32
        l.rtzh_list_hit.argtypes = [C.POINTER(B.rtz_sphere), u64, B.D3, B.D3, f64, f64, C.POINTER(B.rtz_hit)]
        _lib = l
    return _lib


def generate_world(seed: int | None):
    """Scene.init(seed).generateWorld() -> (rtz_sphere array, n).  485 spheres for 0xdeadbeef."""
    buf = (B.rtz_sphere * 512)()
    n = hostlib().rtzh_scene_generate_world(seed or 0, 0 if seed is None else 1, buf, 512)
    return buf, int(n)


def generate_sweep(seed: int, n: int):
    """BASELINE config 5: the final scene generalised to exactly n spheres (16 ... 4096)."""
    buf = (B.rtz_sphere * n)()
    got = hostlib().rtzh_scene_generate_sweep(seed, n, buf)
    if got != n:
        raise ValueError(f"cannot build a {n}-sphere scene")
    return buf, n


def generate_chapter13():
    buf = (B.rtz_sphere * 5)()
    n = hostlib().rtzh_scene_generate_chapter13(buf, 5)
    return buf, int(n)


def camera_build(width, aspect, look_from, look_at, vfov, *, vup=(0, 1, 0), focus_dist=None, defocus_angle=0.0,
                 spp=100, bounce_max=50, seed=None) -> B.rtz_camera:
    cam = B.rtz_camera()
    d3 = lambda v: B.D3(*[float(x) for x in v])
    B.check(hostlib().rtzh_camera_build(width, aspect, d3(look_from), d3(look_at), d3(vup), vfov,
                                        -1.0 if focus_dist is None else focus_dist, defocus_angle, spp, bounce_max,
                                        seed or 0, 0 if seed is None else 1, C.byref(cam)))
    return cam


def main_camera(width: int, spp: int, seed=None) -> B.rtz_camera:
    """The camera of reference src/main.zig:23-31."""
    return camera_build(width, 16.0 / 9.0, (13, 2, 3), (0, 0, 0), 20, focus_dist=10.0, defocus_angle=0.6, spp=spp, seed=seed)


def run_main(img_width: int, spp: int, file_name: str, seed=None, num_gpus: int = 1):
    """main(): renders to images/<file_name> under the current directory (-DnumGpus=N: on N GPUs of the box,
    0 = all, from this one process).  Returns rtz_stats."""
    st = B.rtz_stats()
    B.check(hostlib().rtzh_main(img_width, spp, file_name.encode(), seed or 0, 0 if seed is None else 1, int(num_gpus),
                                C.byref(st)))
    return st


def ppm_save(path: str, width: int, height: int, pixels=None, binary: bool = False) -> None:
    """PPM.init(width, height) [+ caller-filled pixels, 3 f64 each] then PPM.save (P3) / PPM.saveBinary (P6)."""
    arr = None
    if pixels is not None:
        flat = [float(v) for px in pixels for v in px]
        arr = (C.c_double * len(flat))(*flat)
    B.check(hostlib().rtzh_ppm_save(str(path).encode(), width, height, arr, 1 if binary else 0))


def color_from_value(value: int):
    out = B.D3()
    hostlib().rtzh_color_from_value(value, out)
    return tuple(out)


def color_from_rgb(r: int, g: int, b: int):
    out = B.D3()
    hostlib().rtzh_color_from_rgb(r, g, b, out)
    return tuple(out)


def color_to_value(rgb) -> int:
    out = C.c_uint32()
    B.check(hostlib().rtzh_color_to_value(B.D3(*[float(x) for x in rgb]), C.byref(out)))
    return int(out.value)


def color_to_rgb(rgb):
    out = (C.c_uint8 * 3)()
    B.check(hostlib().rtzh_color_to_rgb(B.D3(*[float(x) for x in rgb]), out))
    return tuple(out)
