import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import ctypes
        from svs_b200 import _lib
        lib = _lib.load()
        h = ctypes.c_void_p()
        rc = lib.svsb_create(None, 0, ctypes.byref(h))
        if rc == 0:
            lib.svsb_destroy(h)
            return True
        return False
    except Exception:
        return False


_GPU = None


def pytest_collection_modifyitems(config, items):
    global _GPU
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    if _GPU is None:
        _GPU = _has_gpu()
    if not _GPU:
        skip = pytest.mark.skip(reason="no usable B200 / libsvsb200.so (GPU tests run under gpurun)")
        for it in gpu_items:
            it.add_marker(skip)
