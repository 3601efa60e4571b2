"""N>1 on real GPUs: one process per GPU over NCCL, interleaved tiles gathered to rank 0.  The image
must equal the single-GPU image byte for byte (integer accumulation + counter-based RNG).  Skipped when
the box has a single GPU (the driver's `pytest -m gpu` box); the same logic runs on CPU with gloo in
test_distributed_gloo.py and on one GPU, shard by shard, in test_gpu_parity.py."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

WORKER = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RTZ_ROOT"])
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
pkg = importlib.import_module("raytracing-with-zig_b200")
host = importlib.import_module("raytracing-with-zig_b200.host_api")
sp, n = host.generate_world(0xDEADBEEF)
cam = host.main_camera(320, 12, seed=0xDEADBEEF)
r = pkg.Renderer(lr)
r.upload(sp, n)
img, st = pkg.render_sharded(r, cam, (16, 16))
tot = torch.tensor([float(st.segments)], dtype=torch.float64, device=f"cuda:{lr}")
dist.all_reduce(tot)
if rank == 0:
    whole, wst = r.render(cam)
    assert torch.equal(img, whole), "N-GPU image differs from the 1-GPU image"
    assert int(tot.item()) == wst.segments
    np.save(os.environ["RTZ_OUT"], img.cpu().numpy())
dist.barrier()
dist.destroy_process_group()
'''


def test_nccl_tile_gather_equals_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "img.npy"
    env = dict(os.environ, RTZ_ROOT=str(ROOT), RTZ_OUT=str(out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    img = np.load(out)
    assert img.shape == (180, 320, 3) and img.std() > 10


# ---- N GPUs behind the C ABI: ONE process, no launcher (rtz_multi_*, rtz_render_multi) -------------------------
def _final(pkg, width=320, spp=12):
    import importlib
    host = importlib.import_module("raytracing-with-zig_b200.host_api")
    sp, n = host.generate_world(0xDEADBEEF)
    return sp, n, host.main_camera(width, spp, seed=0xDEADBEEF)


def test_c_abi_multi_on_one_device_equals_rtz_render(pkg):
    """rtz_multi with a single device is rtz_render (runs on the driver's 1-GPU box too)."""
    sp, n, cam = _final(pkg)
    one, st1 = pkg.render_host(cam, sp, n)
    m = pkg.MultiRenderer(1)
    m.upload(sp, n)
    img, st = m.render(cam)
    assert np.array_equal(img, one) and (st.samples, st.segments, st.gpus) == (st1.samples, st1.segments, 1)
    img2, st2 = pkg.render_host_multi(cam, sp, n, num_gpus=1)
    assert np.array_equal(img2, one)
    m.close()
    # more devices than the box has is an argument error, not a silent clamp
    import torch
    try:
        pkg.MultiRenderer(torch.cuda.device_count() + 1)
        raise AssertionError("expected RTZ_ERR_BAD_ARG")
    except pkg.RtzError as e:
        assert e.status == 1


@pytest.mark.parametrize("gather", ["p2p", "nccl"])
def test_c_abi_multi_gpu_equals_single_gpu(pkg, gather):
    """Camera.render on every GPU of the box from ONE host process through the C ABI: the image and the work
    counters equal the single-GPU call byte for byte, for both transports, several device counts and tile sizes."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    sp, n, cam = _final(pkg)
    one, st1 = pkg.render_host(cam, sp, n)
    for world in sorted({2, min(ngpu, 3), ngpu}):
        for tile in ((4, 4), (16, 16), (7, 5)):
            m = pkg.MultiRenderer(world, tile=tile, gather=gather)
            assert (m.gpus, m.gather) == (world, gather)
            m.upload(sp, n)
            for _ in range(2):   # twice: buffers are reused from frame to frame
                img, st = m.render(cam)
                assert np.array_equal(img, one), (world, tile, gather)
                assert (st.samples, st.segments, st.depth_capped, st.absorbed) == (st1.samples, st1.segments, st1.depth_capped, st1.absorbed)
                assert st.gpus == world and st.nan_samples == 0
            m.close()
    # the one-shot form, all devices (what the Zig / C host calls with -DnumGpus)
    img, st = pkg.render_host_multi(cam, sp, n, num_gpus=0)
    assert np.array_equal(img, one) and st.gpus == ngpu
    # a device list: rank 0 (where the image lands) need not be device 0
    m = pkg.MultiRenderer(2, devices=[1, 0], gather=gather)
    m.upload(sp, n)
    img, _ = m.render(cam)
    assert np.array_equal(img, one)
    m.close()
    # the deterministic legacy modes stay whole-frame on rank 0
    import rtzlib as R, ctypes as C
    lcam = R.Camera()
    R.oracle().orc_camera_legacy(400, 16.0 / 9.0, R.MODE_LEGACY_NORMAL, C.byref(lcam))
    two = R.sphere_array([R.make_sphere((0, 0, -1), 0.5, 0), R.make_sphere((0, -100.5, -1), 100, 0)])
    lrgb, _ = pkg.render_host_multi(lcam, two, 2, num_gpus=0)
    assert lrgb.tobytes() == R.read_ppm(R.GOLDEN / "chapter6.ppm")[2]
