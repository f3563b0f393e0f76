import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import __graft_entry__ as entry  # noqa: E402

entry.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test, excluded by default (run with --runslow)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def built_lib():
    """libbtfem.so, built if missing (nvcc cross-compiles without a GPU)."""
    entry.build_library()
    from dmri_fem_cloud_b200 import btfem
    return btfem.load_library()


REF_MESH_DIR = "/root/reference/comri/meshes"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_addoption(parser):
    parser.addoption("--runslow", action="store_true", default=False)


def _gpu_usable():
    """A CUDA device the library can open?  (btfem_create(0) fails without one: the product has no CPU fallback.)"""
    try:
        entry.build_library()
        from dmri_fem_cloud_b200 import btfem
        btfem.BTFem(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if not config.getoption("--runslow"):
        skip = pytest.mark.skip(reason="slow: pass --runslow")
        for item in items:
            if "slow" in item.keywords:
                item.add_marker(skip)
    gpu_items = [item for item in items if "gpu" in item.keywords]
    if gpu_items and not _gpu_usable():
        skip_gpu = pytest.mark.skip(reason="no usable CUDA device / libbtfem.so on this box")
        for item in gpu_items:
            item.add_marker(skip_gpu)
