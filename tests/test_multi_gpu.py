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
