import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/liboracle.so), built on demand.  Test infrastructure only."""
    import rtzlib
    return rtzlib.oracle()


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds librtz.so in-tree if it is stale (nvcc cross-compiles without a GPU)."""
    build = importlib.import_module("raytracing-with-zig_b200.build")
    build.build_all()
    return importlib.import_module("raytracing-with-zig_b200")


@pytest.fixture(scope="session")
def gpu(pkg):
    """A Renderer on cuda:0.  Fails (does not skip) when the CUDA path is unusable on a GPU box."""
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a GPU"
    r = pkg.Renderer(0)
    yield r
    r.close()
