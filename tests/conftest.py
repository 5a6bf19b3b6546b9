import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session.  On a machine without nvcc the prebuilt
    libc2ray_b200.so (it travels with the tree) is used as it is and only the oracle is (re)built with g++."""
    import shutil
    import subprocess
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "c2-ray3dm1d_helium_b200", "libc2ray_b200.so")
    have_nvcc = shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")
    if have_nvcc or not os.path.exists(lib):
        g.build()
    else:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])

sys.path.insert(0, os.path.join(ROOT, "tools"))
