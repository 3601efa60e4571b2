"""Resident renderer: scene and frame stay in HBM; PyTorch supplies device memory and the stream.

`Renderer.render` is the Python face of `Camera.render` (reference src/camera.zig:123-145) on one
GPU; the arithmetic happens in csrc/ (hand-written sm_100a kernels) behind include/rtz.h.
"""
from __future__ import annotations

import ctypes as C

from . import binding as B


def _as_sphere_array(spheres):
    if isinstance(spheres, C.Array):
        return spheres, len(spheres)
    arr = (B.rtz_sphere * len(spheres))()
    for i, s in enumerate(spheres):
        C.memmove(C.byref(arr, i * C.sizeof(B.rtz_sphere)), C.byref(s), C.sizeof(B.rtz_sphere))
    return arr, len(spheres)


def _cast_camera(cam) -> B.rtz_camera:
    """Accept any ctypes struct with the rtz_camera layout (tests use their own mirror class)."""
    if isinstance(cam, B.rtz_camera):
        return cam
    assert C.sizeof(cam) == C.sizeof(B.rtz_camera)
    out = B.rtz_camera()
    C.memmove(C.byref(out), C.byref(cam), C.sizeof(out))
    return out


class Renderer:
    """One rtz_context bound to a CUDA device and to a torch stream of its own on it.

    Stream ordering: the library launches everything on `self.stream` (a real `torch.cuda.Stream`, so its
    handle is never the NULL that `rtz_context_create` reads as "make your own stream").  Every call first
    makes that stream wait for torch's current stream (`_order_after_torch`), so tensors produced by torch or
    by a collective (the NCCL tile gather makes the current stream wait for its result) are complete before
    the library reads them; every library call blocks until its own work is done, so results are complete
    for whatever torch does next."""

    def __init__(self, device: int | None = None):
        import torch

        if not torch.cuda.is_available():
            raise B.RtzError(2, "no CUDA device: raytracing-with-zig_b200 has no CPU fallback")
        self.torch = torch
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.lib = B.lib()
        self.stream = torch.cuda.Stream(device=self.device)
        assert self.stream.cuda_stream != 0
        self._ctx = C.c_void_p()
        B.check(self.lib.rtz_context_create(self.device, C.c_void_p(self.stream.cuda_stream), C.byref(self._ctx)))
        self.n_spheres = 0

    def _order_after_torch(self):
        self.stream.wait_stream(self.torch.cuda.current_stream(self.device))

    def close(self):
        if self._ctx:
            self.lib.rtz_context_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, spheres, n: int | None = None):
        """HittableList.add for the whole list: flatten to the device SoA and copy to HBM."""
        arr, cnt = _as_sphere_array(spheres)
        n = cnt if n is None else int(n)
        ptr = C.cast(arr, C.POINTER(B.rtz_sphere))
        B.check(self.lib.rtz_scene_upload(self._ctx, ptr, n))
        self.n_spheres = n

    def generate(self, kind: int, seed: int, n: int = 0):
        """Scene.init(seed) + generateWorld / generateChapter13 / the config-5 sweep scene, generated on the
        device with the reference's PRNG stream and installed as this renderer's scene.
        Returns (spheres, count, prng_state): the f64 spheres in list order and Scene.prng after generation."""
        cap = int(n) if kind == B.SCENE_SWEEP else 22 * 22 + 4
        arr = (B.rtz_sphere * max(cap, 1))()
        cnt = C.c_uint64(0)
        state = (C.c_uint64 * 4)()
        B.check(self.lib.rtz_scene_generate(self._ctx, int(kind), C.c_uint64(seed), C.c_uint64(n), arr, cap,
                                            C.byref(cnt), state))
        self.n_spheres = int(cnt.value)
        return arr, int(cnt.value), [int(x) for x in state]

    def shard_pixels(self, width: int, height: int, shard: B.rtz_shard | None) -> int:
        return int(self.lib.rtz_shard_pixels(width, height, C.byref(shard) if shard is not None else None))

    def render(self, camera, shard: B.rtz_shard | None = None, out=None):
        """Render the frame (or this rank's tiles) into a device uint8 tensor; returns (tensor, stats)."""
        torch = self.torch
        cam = _cast_camera(camera)
        if shard is None:
            shape = (int(cam.height), int(cam.width), 3)
        else:
            shape = (self.shard_pixels(cam.width, cam.height, shard), 3)
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=f"cuda:{self.device}")
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.uint8 and out.numel() == shape[0] * shape[1] * (shape[2] if len(shape) == 3 else 1)
        st = B.rtz_stats()
        self._order_after_torch()
        B.check(self.lib.rtz_render_resident(self._ctx, C.byref(cam), C.byref(shard) if shard is not None else None,
                                             C.c_void_p(out.data_ptr()), C.byref(st)))
        return out, st

    def deinterleave(self, gathered, width: int, height: int, world: int, tile_w: int, tile_h: int, out=None):
        torch = self.torch
        if out is None:
            out = torch.empty((height, width, 3), dtype=torch.uint8, device=gathered.device)
        self._order_after_torch()   # `gathered` may still be in flight on torch's / NCCL's streams
        B.check(self.lib.rtz_deinterleave(self._ctx, width, height, world, tile_w, tile_h,
                                          C.c_void_p(gathered.data_ptr()), C.c_void_p(out.data_ptr())))
        return out


class MultiRenderer:
    """`rtz_multi`: ONE process driving N GPUs of the box behind the C ABI (no torchrun, no torch.distributed).

    The frame is cut into interleaved tiles, every device traces its tiles, and the bytes reach device 0 either
    through the fused resolve + peer-memory store (`gather="p2p"`) or a grouped ncclSend/ncclRecv (`"nccl"`)."""

    def __init__(self, num_gpus: int = 0, devices=None, tile=(4, 4), gather: str = "auto"):
        self.lib = B.lib()
        self._m = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices) if devices else None
        g = {"auto": B.GATHER_AUTO, "p2p": B.GATHER_P2P, "nccl": B.GATHER_NCCL}[gather]
        B.check(self.lib.rtz_multi_create(int(num_gpus), devs, int(tile[0]), int(tile[1]), g, C.byref(self._m)))
        self.gpus = int(self.lib.rtz_multi_gpus(self._m))
        self.gather = {B.GATHER_P2P: "p2p", B.GATHER_NCCL: "nccl"}[int(self.lib.rtz_multi_gather(self._m))]

    def close(self):
        if self._m:
            self.lib.rtz_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, spheres, n: int | None = None):
        arr, cnt = _as_sphere_array(spheres)
        n = cnt if n is None else int(n)
        B.check(self.lib.rtz_multi_scene_upload(self._m, C.cast(arr, C.POINTER(B.rtz_sphere)), n))

    def render(self, camera, to_host: bool = True):
        """Returns (numpy uint8 [H,W,3] or None, stats); the image also stays resident on device 0."""
        import numpy as np

        cam = _cast_camera(camera)
        H, W = int(cam.height), int(cam.width)
        rgb = np.empty((H, W, 3), dtype=np.uint8) if to_host else None
        st = B.rtz_stats()
        ptr = rgb.ctypes.data_as(C.POINTER(C.c_uint8)) if to_host else None
        B.check(self.lib.rtz_multi_render(self._m, C.byref(cam), ptr, None, C.byref(st)))
        return rgb, st


def render_host_multi(camera, spheres, n: int | None = None, num_gpus: int = 0):
    """`rtz_render_multi`: the one-shot C-ABI call with HOST buffers on `num_gpus` devices (0 = all)."""
    import numpy as np

    l = B.lib()
    cam = _cast_camera(camera)
    arr, cnt = _as_sphere_array(spheres)
    n = cnt if n is None else int(n)
    rgb = np.empty((int(cam.height), int(cam.width), 3), dtype=np.uint8)
    st = B.rtz_stats()
    B.check(l.rtz_render_multi(C.byref(cam), C.cast(arr, C.POINTER(B.rtz_sphere)), n, int(num_gpus),
                               rgb.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st)))
    return rgb, st


def render_host(camera, spheres, n: int | None = None, want_linear: bool = False):
    """The one-shot C-ABI call with HOST buffers (`rtz_render`): what a Zig/C host would call.

    Returns (numpy uint8 [H,W,3], stats[, numpy float64 [H,W,3]]).  Host->device copy of the scene
    and device->host copy of the image happen inside the call."""
    import numpy as np

    l = B.lib()
    cam = _cast_camera(camera)
    arr, cnt = _as_sphere_array(spheres)
    n = cnt if n is None else int(n)
    H, W = int(cam.height), int(cam.width)
    rgb = np.empty((H, W, 3), dtype=np.uint8)
    st = B.rtz_stats()
    sp = C.cast(arr, C.POINTER(B.rtz_sphere))
    if want_linear:
        lin = np.empty((H, W, 3), dtype=np.float64)
        B.check(l.rtz_render_linear(C.byref(cam), sp, n, rgb.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    lin.ctypes.data_as(C.POINTER(C.c_double)), C.byref(st)))
        return rgb, st, lin
    B.check(l.rtz_render(C.byref(cam), sp, n, rgb.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st)))
    return rgb, st
