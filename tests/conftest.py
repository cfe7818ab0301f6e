import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def _gpu_probe():
    """One throw-away handle: is libpfb200.so built and is there a usable CUDA device?  Returns the
    reason the GPU tier cannot run, or None."""
    try:
        import ctypes
        from qnmh_sysid2018_b200 import _lib
        lib = _lib.load()
        h = ctypes.c_void_p()
        rc = lib.pfb200_create(0, 0, 0, ctypes.byref(h))
        if rc != 0:
            return "pfb200_create failed: " + lib.pfb200_last_error().decode("utf-8", "replace")
        lib.pfb200_destroy(h)
        return None
    except Exception as err:  # library not built, ...
        return "libpfb200 unavailable: %s" % err


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/python/state")
    skip_ref = pytest.mark.skip(reason="/root/reference not mounted")
    gpu_items = [item for item in items if "gpu" in item.keywords]
    skip_gpu = None
    # a host WITH an NVIDIA device never skips: there a library that does not load must fail loudly
    if gpu_items and not os.path.exists("/dev/nvidiactl"):
        why = _gpu_probe()
        if why is not None:
            skip_gpu = pytest.mark.skip(reason="GPU tier: " + why)
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if skip_gpu is not None and "gpu" in item.keywords:
            item.add_marker(skip_gpu)
