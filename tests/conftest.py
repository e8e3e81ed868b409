import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pcamv():
    import pcamv_loader
    return pcamv_loader.load()


@pytest.fixture(scope="session")
def cuda_lib(pcamv):
    """libpcamv_cuda.so, built in-tree (nvcc cross-compiles without a GPU)."""
    pcamv.build.build_cuda()
    return pcamv.load_library()
