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


def test_shim_compiles_against_the_reference_headers(tmp_path):
    """integration/CudaWrappers.h - the file a maintainer of the reference drops over src/cuda/CudaWrappers.h - compiled
    against the reference's OWN src/AppParams.h and src/cuda/Mat.h: the 13 wrapper signatures of
    src/cuda/CudaWrappers.h:22-36 type-check with the reference's types and the reinterpret_cast layout claims hold
    (static_asserts in the shim).  Needs the reference checkout (dev container only)."""
    import shutil
    import subprocess
    ref = "/root/reference/src"
    if not os.path.isdir(ref) or shutil.which("g++") is None:
        pytest.skip("reference checkout or g++ not available")
    tree = tmp_path / "cuda"
    tree.mkdir()
    shutil.copy(os.path.join(ROOT, "integration", "CudaWrappers.h"), tree / "CudaWrappers.h")
    tu = tmp_path / "tu.cpp"
    tu.write_text('''#include <iostream>
#include <string>
#include <vector>
using namespace std;
#include "cuda/CudaWrappers.h"
hkf_ctx* g_hkf = nullptr;
void touch()
{
    Mat44 m = Mat44::getIdentity();
    CameraParams c = AppParams::instance()->_depth_camera_params;
    cudaTruncDepth(0.3f, 4.f); cudaBiliearFilterDepth(2.f, 0.02f); cudaCalculateNewVertices(c); cudaCalculateNewNormals();
    cudaDownSampleNewVertices(); cudaDownSampleNewNormals(); cudaDownSampleModelVertices(); cudaDownSampleModelNormals();
    cudaCalPointToPlaneErrSolverParams(0, m, m, c, 0.1f, 0.2f); cudaCalSDFSolverParams(c, m);
    cudaIntegrateVolume(true, false, m, AppParams::instance()->_integrate_params, c, AppParams::instance()->_rgb_camera_params);
    cudaRaycastingVolume(true, m, AppParams::instance()->_raycast_params, c, 0.3f, 4.f);
    cudaMarchingcube(true, 1.f);
}
''')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", f"-I{tmp_path}", f"-I{os.path.join(ROOT, 'include')}", f"-I{ref}", f"-I{ref}/cuda",
                        "-I/usr/local/cuda/include", str(tu)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
