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


def pytest_collection_modifyitems(config, items):
    if config.getoption("--runslow"):
        return
    skip = pytest.mark.skip(reason="slow: pass --runslow")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)
