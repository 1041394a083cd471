import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_count():
    """Number of CUDA devices seen by the product library's runtime (0 without a GPU or without the .so:
    the library has no CPU fallback, so the `gpu` tests cannot run there and are skipped, not failed)."""
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        if rt.cudaGetDeviceCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 0


def pytest_collection_modifyitems(config, items):
    if not any("gpu" in it.keywords for it in items):
        return
    n = _gpu_count()
    if n > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (tnqs_b200 has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
