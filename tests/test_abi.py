"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol include/hkf.h
declares, has the POD layouts the reference passes by value, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hybkinectfu_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hkf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hkf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/hkf.h but not exported"
    assert sorted(capi.EXPORTS) == syms, "capi.EXPORTS is out of sync with include/hkf.h"


def test_pod_layouts_match_reference_sizes():
    # sizes verified against the reference with sizeof: CameraParams 24, Mat44 64,
    # IntegrateParams 8, RayCasterParams 4, Triangle 72
    assert C.sizeof(capi.CameraParams) == 24
    assert C.sizeof(capi.Mat44) == 64
    assert C.sizeof(capi.IntegrateParams) == 8
    assert C.sizeof(capi.RayCasterParams) == 4


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.HkfError) as e:
        capi.Context(capi.Config(cols=64, rows=48, resolution=32))
    assert "no CPU fallback" in str(e.value)


def test_config_defaults_follow_reference_config_ini():
    cfg = capi.Config()
    # root config.ini:17-48 (depth camera 640x480 @525, S=2, wmax=128, trunc 0.04, depth window 0.3..4)
    assert (cfg.cols, cfg.rows, cfg.fx, cfg.cx) == (640, 480, 525.0, 319.5)
    assert cfg.volume_size == 2.0 and cfg.weight_max == 128.0 and cfg.sdf_trunc == 0.04
    # fRayIncrement = factor * trunc in float (src/AppParamsProducer.cpp:113-117)
    assert np.float32(cfg.ray_increment) == np.float32(0.8) * np.float32(0.04)
    c1 = cfg.level_camera(1)
    assert (c1.cols, c1.rows) == (320, 240) and c1.fx == 262.5 and c1.cx == 159.75
