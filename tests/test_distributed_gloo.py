"""N>1 host logic on CPU: two gloo ranks shard a frame by interleaved tiles, gather to rank 0 with
the package's own gather + index map, and the result must equal the single-rank frame byte for
byte.  The per-rank pixels come from the oracle's device mirror (no GPU here)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rtzlib as R

ROOT = Path(__file__).resolve().parent.parent
TILE = (16, 16)


def _worker(rank, world, port, out_path):
    import importlib
    sys.path.insert(0, str(ROOT)), sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("raytracing-with-zig_b200")
    orc = R.oracle()
    sp, n = R.chapter13_scene()
    cam = R.build_camera(100, 16.0 / 9.0, (-2, 2, 1), (0, 0, -1), 20, spp=4, seed=9)
    per_rank, idx = pkg.tile_index_map(cam.width, cam.height, world, *TILE)
    sh = R.Shard(rank, world, *TILE)
    local = np.zeros((per_rank, 3), np.uint8)
    assert orc.orc_render_mirror(C.byref(cam), sp, n, 9, 2, C.byref(sh), local.ctypes.data_as(C.POINTER(C.c_uint8)), None,
                                 None) == 0
    from importlib import import_module
    gathered = import_module("raytracing-with-zig_b200.distributed").gather_tiles(torch.from_numpy(local), world, rank)
    if rank == 0:
        img = gathered.numpy()[idx.reshape(-1)].reshape(cam.height, cam.width, 3)
        np.save(out_path, img)
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_tile_gather_equals_single_rank(orc, tmp_path, world):
    out = str(tmp_path / "img.npy")
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    img = np.load(out)
    sp, n = R.chapter13_scene()
    cam = R.build_camera(100, 16.0 / 9.0, (-2, 2, 1), (0, 0, -1), 20, spp=4, seed=9)
    whole = np.zeros((cam.height, cam.width, 3), np.uint8)
    assert orc.orc_render_mirror(C.byref(cam), sp, n, 9, 2, None, whole.ctypes.data_as(C.POINTER(C.c_uint8)), None, None) == 0
    assert np.array_equal(img, whole)


def test_tile_index_map_is_a_bijection_onto_valid_slots():
    import importlib
    pkg = importlib.import_module("raytracing-with-zig_b200")
    for (w, h, world, tw, th) in [(100, 56, 2, 16, 16), (37, 29, 3, 7, 5), (1200, 675, 8, 16, 16), (5, 5, 4, 8, 8)]:
        per_rank, idx = pkg.tile_index_map(w, h, world, tw, th)
        flat = idx.reshape(-1)
        assert len(np.unique(flat)) == w * h and flat.min() >= 0 and flat.max() < world * per_rank
        # rank sizes are balanced to within one tile
        counts = np.bincount(flat // per_rank, minlength=world)
        assert counts.max() - counts.min() <= tw * th * max(1, (h + th - 1) // th)
