import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _cuda_device_count():
    """devices the CUDA runtime can see, without importing torch (0 when the driver is missing)"""
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
        except OSError:
            return 0
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def pytest_collection_modifyitems(config, items):
    """plain `pytest tests` on a box without a usable GPU skips the gpu-marked tests instead of failing in hkf_create"""
    if not any("gpu" in it.keywords for it in items) or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (the library has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def stages():
    return dict(np.load(os.path.join(GOLDEN, "stages_small.npz")))


@pytest.fixture(scope="session")
def pipeline():
    return dict(np.load(os.path.join(GOLDEN, "pipeline_small.npz")))


def small_config(resolution=64, levels=3):
    from hybkinectfu_b200 import synth
    from hybkinectfu_b200.capi import Config
    cam = synth.Camera().scaled(192, 144)
    return Config(cols=192, rows=144, cx=cam.cx, cy=cam.cy, fx=cam.fx, fy=cam.fy, resolution=resolution,
                  pyramid_levels=levels, max_triangles=400000)


def bits_equal(a, b):
    """float arrays equal bit for bit, treating +0 and -0 as equal and NaN == NaN"""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if a.shape != b.shape:
        return False
    same = (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)) | (np.isnan(a) & np.isnan(b))
    return bool(same.all())


def n_mismatch(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)) | (np.isnan(a) & np.isnan(b))
    return int((~same).sum())


def canon_tris(t):
    """order-independent form of a triangle soup: rows of 9 position floats, lexicographically sorted"""
    if len(t) == 0:
        return np.zeros((0, 9), np.float32)
    p = np.ascontiguousarray(np.asarray(t)[:, :, :3]).reshape(len(t), 9)
    key = np.lexsort(p.T[::-1])
    return p[key]
